"""train -- drop-in for /root/reference/tinyfaces/trainer.py:68-90 (the step loop), plus what the reference does not
have: the data-parallel step (one process per GPU, batch-sharded, gradients SUM-reduced over NCCL -- the reference loss
is a sum, loss.py:87-88, so the reduction must not divide), the overlapped bucket pipeline (optim.reduce_and_step) and
a CUDA-graph replay of the whole step (GraphedTrainStep: ~900 launches per step collapse into one graph launch).
"""
import torch
import torch.distributed as dist

from .optim import FlatSGD, reduce_and_step


def print_state(idx, epoch, size, loss_cls, loss_reg):
    head = "Epoch: [{0}][{1}/{2}]\t".format(epoch, idx, size) if epoch >= 0 else "Val: [{0}/{1}]\t".format(idx, size)
    print(head + "\tloss_cls: {0:.6f}\tloss_reg: {1:.6f}".format(loss_cls, loss_reg))


def allreduce_gradients(parameters, group=None):
    """SUM all-reduce of every existing .grad as ONE flat fp32 buffer (fallback for a plain torch optimizer; with FlatSGD
    the gradients already live in one buffer and are reduced bucket by bucket DURING the backward)."""
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
    torch._foreach_copy_(grads, views)        # one multi-tensor launch instead of ~300 small copies


def train_step(model, loss_fn, optimizer, img, class_map, regression_map, group=None):
    """One forward / loss / backward / all-reduce / SGD step on tensors already on the device (trainer.py:78-87).
    With a FlatSGD optimizer the all-reduce and the optimizer run per gradient bucket on a communication stream while the
    backward is still producing the earlier layers' gradients."""
    output = model(img)
    loss = loss_fn(output, class_map, regression_map)
    optimizer.zero_grad()
    loss.backward()
    if isinstance(optimizer, FlatSGD):
        reduce_and_step(optimizer, group)
    else:
        allreduce_gradients([p for g in optimizer.param_groups for p in g["params"]], group)
        optimizer.step()
    return loss


def train_step_flat(model, loss_fn, optimizer, img, class_map, regression_map, group=None):
    """The same step without the autograd engine (FlatSGD only): the loss kernel already returns d loss / d output, which
    goes straight into the executor's backward.  This is what GraphedTrainStep captures (the engine's end-of-backward
    stream bookkeeping waits on events recorded outside the capture, which a stream capture rejects)."""
    output = model.forward_train_flat(img)
    loss, grad = loss_fn.loss_and_grad(output, class_map, regression_map)
    model.backward_flat(grad)
    reduce_and_step(optimizer, group)
    return loss


class GraphedTrainStep:
    """The whole training step (forward, loss incl. OHEM + device sampler, backward, bucketed all-reduce, SGD) captured
    ONCE into a CUDA graph and replayed per batch: the ~900 kernel launches (and their ~2 ms of host + launch-gap time,
    which is most of a step at the small per-GPU batch of BASELINE configs[3]) become one graph launch.

    Requirements: static shapes, ``DetectionCriterion(sampler="device")``, a FlatSGD optimizer whose schedule lives on the
    device (``optimizer.steplr``), and ``warmup`` eager steps first (they size the workspace and fill the library's caches;
    they are real optimizer steps on the given batch)."""

    def __init__(self, model, loss_fn, optimizer, img, class_map, regression_map, group=None, warmup=2):
        if not isinstance(optimizer, FlatSGD):
            raise RuntimeError("GraphedTrainStep needs a tinyfaces_b200.optim.FlatSGD optimizer")
        if getattr(loss_fn, "sampler", "device") != "device":
            raise RuntimeError("GraphedTrainStep needs the device sampler (the numpy sampler runs on the host)")
        self.model, self.loss_fn, self.optimizer, self.group = model, loss_fn, optimizer, group
        self.static = tuple(t.detach().clone().float().contiguous() for t in (img, class_map, regression_map))
        self.batch = img.shape[0]
        dev = img.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.static[1].copy_(class_map)                      # OHEM edits the class map in place
                train_step_flat(model, loss_fn, optimizer, *self.static, group=group)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        self.static[1].copy_(class_map)
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = train_step_flat(model, loss_fn, optimizer, *self.static, group=group)
        # the capture itself did not execute anything: undo its host-side meter bookkeeping
        for m in (loss_fn.class_average, loss_fn.reg_average):
            m.num_averaged -= self.batch

    def __call__(self, img, class_map, regression_map):
        """Copies the batch into the graph's static inputs (device->device, or host->device for pinned host tensors),
        replays, returns the (static) loss tensor -- valid until the next call."""
        for dst, src in zip(self.static, (img, class_map, regression_map)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for m in (self.loss_fn.class_average, self.loss_fn.reg_average):
            m.num_averaged += self.batch
        return self.loss


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


def train_pipelined(model, loss_fn, optimizer, batches, device, group=None, with_meters=False, graph=False):
    """Step loop over an iterable of HOST batches with the copies and the loss read-back taken off the critical path:
    batch i+1 is staged while batch i computes, and the loss of step i is read (from a pinned buffer) only after step
    i+1 has been enqueued.  Yields one Python float per step, in order; with_meters=True yields
    (loss, class_average, reg_average) -- the running averages AS OF THAT STEP, snapshotted on the device right after it
    (reading loss_fn.class_average.average from the consumer would see step i+1's update and synchronise on it).
    graph=True replays a GraphedTrainStep (captured on the first batch; its warm-up steps train on that batch)."""
    pipe = InputPipeline(device)
    host_loss = torch.empty((2, 3), dtype=torch.float32).pin_memory()
    loss_ready = [None, None]
    it = iter(batches)
    nxt = next(it, None)
    if nxt is not None:
        pipe.stage(*nxt)
    i = 0
    pending = None                                   # index of the step whose loss has not been yielded yet
    graphed = graph if isinstance(graph, GraphedTrainStep) else None
    while nxt is not None:
        x, c, r = pipe.next()
        if graph and graphed is None:
            graphed = GraphedTrainStep(model, loss_fn, optimizer, x, c, r, group)
        loss = graphed(x, c, r) if graphed is not None else train_step(model, loss_fn, optimizer, x, c, r, group)
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
