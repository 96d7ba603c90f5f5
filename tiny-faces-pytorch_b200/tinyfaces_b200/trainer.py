"""train -- drop-in for /root/reference/tinyfaces/trainer.py:68-90 (the step loop), plus the data-parallel
pieces the reference does not have: one process per GPU, batch-sharded, gradients SUM-reduced over NCCL
(the reference loss is a sum, loss.py:87-88, so the reduction must not divide).
"""
import torch
import torch.distributed as dist


def print_state(idx, epoch, size, loss_cls, loss_reg):
    head = "Epoch: [{0}][{1}/{2}]\t".format(epoch, idx, size) if epoch >= 0 else "Val: [{0}/{1}]\t".format(idx, size)
    print(head + "\tloss_cls: {0:.6f}\tloss_reg: {1:.6f}".format(loss_cls, loss_reg))


def allreduce_gradients(parameters, group=None):
    """SUM all-reduce of every existing .grad as ONE flat fp32 buffer (110.9 MB for the trunk + heads)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    torch._foreach_copy_(grads, views)        # one multi-tensor launch instead of ~300 small copies (1 ms per step at N > 1)


def train_step(model, loss_fn, optimizer, img, class_map, regression_map, group=None):
    """One forward / loss / backward / (all-reduce) / SGD step on tensors already on the device."""
    output = model(img)
    loss = loss_fn(output, class_map, regression_map)
    optimizer.zero_grad()
    loss.backward()
    allreduce_gradients([p for g in optimizer.param_groups for p in g["params"]], group)
    optimizer.step()
    return loss


class InputPipeline:
    """Double-buffered host->device staging of (img, class_map, regression_map) batches on a copy stream, so that the
    H2D transfer of batch i+1 overlaps the compute of batch i (the reference copies synchronously in the step loop,
    trainer.py:73-76).  Device slots are allocated once; ordering is by events, never by host synchronisation."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.slots = [None] * depth
        self.ready = [None] * depth          # recorded on the copy stream when a slot's copies are enqueued
        self.free = [None] * depth           # recorded on the compute stream when a slot's batch has been consumed
        self.head = self.tail = 0            # staged batches are slots [tail, head)

    def stage(self, img, class_map, regression_map):
        """Start the asynchronous copy of one host batch (pinned memory for true overlap) into the next slot."""
        assert self.head - self.tail < self.depth, "InputPipeline: all slots are staged"
        k = self.head % self.depth
        host = (img, class_map, regression_map)
        if self.slots[k] is None or any(d.shape != h.shape for d, h in zip(self.slots[k], host)):
            self.slots[k] = tuple(torch.empty(h.shape, dtype=torch.float32, device=self.device) for h in host)
            self.free[k] = None
        with torch.cuda.stream(self.copy_stream):
            if self.free[k] is not None:
                self.copy_stream.wait_event(self.free[k])            # the step that used this slot has finished with it
            else:
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            for d, h in zip(self.slots[k], host):
                d.copy_(h, non_blocking=True)                        # (also converts uint8 / float64 sources to float32)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ready[k] = ev
        self.head += 1

    def next(self):
        """Device tensors of the oldest staged batch; the current stream waits for its copies."""
        assert self.head > self.tail, "InputPipeline: nothing staged"
        k = self.tail % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        return self.slots[k]

    def release(self):
        """Call after the step that consumed next()'s tensors has been enqueued."""
        k = self.tail % self.depth
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[k] = ev
        self.tail += 1


def _meter_value(meter, device):
    a = getattr(meter, "_average", 0)
    return a.detach().to(torch.float32) if isinstance(a, torch.Tensor) else torch.tensor(float(a), device=device)


def train_pipelined(model, loss_fn, optimizer, batches, device, group=None, with_meters=False):
    """Step loop over an iterable of HOST batches with the copies and the loss read-back taken off the critical path:
    batch i+1 is staged while batch i computes, and the loss of step i is read (from a pinned buffer) only after step
    i+1 has been enqueued.  Yields one Python float per step, in order; with_meters=True yields
    (loss, class_average, reg_average) -- the running averages AS OF THAT STEP, snapshotted on the device right after it
    (reading loss_fn.class_average.average from the consumer would see step i+1's update and synchronise on it)."""
    pipe = InputPipeline(device)
    host_loss = torch.empty((2, 3), dtype=torch.float32).pin_memory()
    loss_ready = [None, None]
    it = iter(batches)
    nxt = next(it, None)
    if nxt is not None:
        pipe.stage(*nxt)
    i = 0
    pending = None                                   # index of the step whose loss has not been yielded yet
    while nxt is not None:
        x, c, r = pipe.next()
        loss = train_step(model, loss_fn, optimizer, x, c, r, group)
        pipe.release()
        if with_meters:
            snap = torch.stack([loss.detach().to(torch.float32), _meter_value(loss_fn.class_average, pipe.device),
                                _meter_value(loss_fn.reg_average, pipe.device)])
            host_loss[i % 2].copy_(snap, non_blocking=True)
        else:
            host_loss[i % 2, 0].copy_(loss.detach(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(pipe.device))
        loss_ready[i % 2] = ev
        nxt = next(it, None)
        if nxt is not None:
            pipe.stage(*nxt)                         # overlaps the step just enqueued
        if pending is not None:
            loss_ready[pending % 2].synchronize()
            yield _emit(host_loss[pending % 2], with_meters)
        pending = i
        i += 1
    if pending is not None:
        loss_ready[pending % 2].synchronize()
        yield _emit(host_loss[pending % 2], with_meters)


def _emit(row, with_meters):
    return (float(row[0]), float(row[1]), float(row[2])) if with_meters else float(row[0])


def train(model, loss_fn, optimizer, dataloader, epoch, device):
    """trainer.py:68-90 (same signature, same per-iteration printout), on the pipelined loop."""
    model = model.to(device)
    model.train()
    n = len(dataloader)
    for idx, (_loss, cls_avg, reg_avg) in enumerate(train_pipelined(model, loss_fn, optimizer, dataloader, device, with_meters=True)):
        print_state(idx, epoch, n, cls_avg, reg_avg)            # the averages as of step idx (trainer.py:89-90)
