"""Flat parameter / gradient storage, the bucketed overlapped gradient all-reduce and the SGD(+StepLR) step that runs as
each bucket's epilogue -- the optimizer side of /root/reference/tinyfaces/main.py:67-70,81-83 and the collective of the
batch-sharded training step (the reference itself is single-device, trainer.py:83-87).

Layout.  Every parameter the executor trains (conv / BN / head tensors; not the unused ``model.fc`` nor the lr-0 upsample
kernel) is re-pointed into ONE fp32 buffer, its gradient into a second one, both laid out in the order in which the
backward COMPLETES them (score_res4, layer3.22, ..., layer1.0, stem), each tensor padded to 4 floats.  A bucket is a
contiguous slice; ``tf_model_backward_ex`` records an event per bucket as soon as its last gradient has been enqueued, so

    comm stream:  wait(event k) -> all_reduce(SUM, grad[bucket k]) -> tf_sgd_step(bucket k)

runs while the backward is still working on the earlier layers: no ``torch.cat``, no copy-back, no separate optimizer pass.
The gradient is SUM-reduced (the reference loss is a sum, loss.py:87-88).
"""
import ctypes
import re

import torch
import torch.distributed as dist

from ._lib import check, lib, stream_ptr

_BLOCKS_PER_LAYER = (3, 4, 23)


def _block_of(name):
    """Forward index of the residual block whose backward completes this parameter's gradient (see tf_model_backward_ex)."""
    m = re.match(r"model\.layer(\d)\.(\d+)\.", name)
    if m:
        return sum(_BLOCKS_PER_LAYER[: int(m.group(1)) - 1]) + int(m.group(2))
    if name.startswith("score_res4"):
        return 29
    if name.startswith("score_res3"):
        return 7
    return -1                                            # stem: model.conv1 / model.bn1


class FlatParams:
    """Re-points the trainable executor parameters of ``model`` and their gradients into flat buffers (call AFTER
    ``model.to(device)``).  ``model.parameters()`` / ``state_dict()`` keep working: the tensors are views."""

    def __init__(self, model, bucket_bytes=24 << 20):
        ex = model._executor
        named = dict(model.named_parameters())
        names = [n for n in ex.names if n in named and named[n].requires_grad and n != "score4_upsample.weight"]
        order = sorted(range(len(names)), key=lambda i: -_block_of(names[i]))          # stable: completion order
        self.names = [names[i] for i in order]
        self.params = [named[n] for n in self.names]
        dev = self.params[0].device
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.offsets, self.total = offs, off
        self.flat_param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad_views = []
        for p, o in zip(self.params, offs):
            v = self.flat_param[o:o + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v
            g = self.flat_grad[o:o + p.numel()].view_as(p)
            p.grad = g
            self.grad_views.append(g)
        model.__dict__.pop("_ptr_table", None)           # the parameter addresses changed
        # ---- buckets: cut at block boundaries once a bucket holds >= bucket_bytes
        self.buckets = []                                # (begin, end, first_block)
        begin, cur_block = 0, None
        for i, (n, o) in enumerate(zip(self.names, offs)):
            b = _block_of(n)
            if cur_block is not None and b != cur_block and (o - begin) * 4 >= bucket_bytes:
                self.buckets.append((begin, o, cur_block))
                begin = o
            cur_block = b
        self.buckets.append((begin, off, cur_block))
        self.events = []
        for _ in self.buckets:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))    # forces the lazy cudaEvent_t into existence
            self.events.append(ev)
        self._ev_ptrs = (ctypes.c_void_p * len(self.events))(*[e.cuda_event for e in self.events])
        self._ev_blocks = (ctypes.c_int * len(self.events))(*[b[2] for b in self.buckets])
        self.device = dev
        self.comm_stream = torch.cuda.Stream(dev)
        model._flat = self

    def is_valid(self):
        """True while every parameter still views its slice of the flat buffer (a .to() / .cuda() to another device or a
        dtype conversion re-allocates parameter storage and breaks that)."""
        base = self.flat_param.data_ptr()
        return all(p.data_ptr() == base + 4 * o for p, o in zip(self.params, self.offsets))

    def grad_pointer_table(self, ex_names):
        """ctypes table of gradient addresses in executor order (NULL for tensors that are not trained)."""
        by_name = dict(zip(self.names, self.grad_views))
        t = (ctypes.c_void_p * len(ex_names))()
        for i, n in enumerate(ex_names):
            g = by_name.get(n)
            t[i] = g.data_ptr() if g is not None else None
        return t


class FlatSGD(torch.optim.Optimizer):
    """torch.optim.SGD(params, lr, momentum, weight_decay) over the parameter groups of
    ``DetectionModel.learnable_parameters`` (model.py:67-87), executed by ``tf_sgd_step`` on the flat buffers.
    ``param_groups`` keeps torch's format (``lr`` per group is read at every step, so ``torch.optim.lr_scheduler.StepLR``
    works unchanged); ``set_lr_scale`` / ``steplr`` keep the schedule on the device instead (CUDA-graph friendly).
    Parameters outside the flat store (``model.fc``: never has a gradient; the upsample kernel: lr 0) are never touched,
    exactly like torch.optim.SGD skips ``grad is None`` / multiplies by lr 0."""

    def __init__(self, model, param_groups, momentum=0.9, weight_decay=0.0, bucket_bytes=24 << 20):
        defaults = dict(lr=0.0, momentum=momentum, weight_decay=weight_decay)
        super().__init__(param_groups, defaults)
        self.flat = getattr(model, "_flat", None) or FlatParams(model, bucket_bytes)
        f = self.flat
        self.momentum_buf = torch.zeros(f.total, dtype=torch.float32, device=f.device)
        gid = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                gid[id(p)] = gi
        self._group_of = [gid[id(p)] for p in f.params]
        self.lr_scale = torch.ones(1, dtype=torch.float32, device=f.device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=f.device)
        self._plans = None

    def zero_grad(self, set_to_none=False):
        """The flat backward OVERWRITES every gradient it owns (= zero_grad() + backward()), so there is nothing to clear."""
        return None

    def _bucket_plan(self):
        """Per bucket: segment table (runs of equal (lr, wd)) as ctypes arrays; rebuilt when a group's lr / wd changes."""
        key = tuple((g["lr"], g["weight_decay"], g["momentum"]) for g in self.param_groups)
        if self._plans is not None and self._plans[0] == key:
            return self._plans[1]
        f = self.flat
        plans = []
        for (b0, b1, _blk) in f.buckets:
            begins, lrs, wds = [], [], []
            for o, gi in zip(f.offsets, self._group_of):
                if o < b0 or o >= b1:
                    continue
                g = self.param_groups[gi]
                if not begins or (lrs[-1], wds[-1]) != (g["lr"], g["weight_decay"]):
                    begins.append(o - b0); lrs.append(g["lr"]); wds.append(g["weight_decay"])
            n = len(begins)
            plans.append((n, (ctypes.c_int64 * n)(*begins), (ctypes.c_float * n)(*lrs), (ctypes.c_float * n)(*wds)))
        self._plans = (key, plans)
        return plans

    def step_bucket(self, k, stream=None):
        """SGD on bucket k, enqueued on `stream` (default: the current one)."""
        f = self.flat
        b0, b1, _ = f.buckets[k]
        n, begins, lrs, wds = self._bucket_plan()[k]
        st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else stream_ptr(f.device)
        check(lib().tf_sgd_step(f.flat_param[b0:b1].data_ptr(), f.flat_grad[b0:b1].data_ptr(), self.momentum_buf[b0:b1].data_ptr(),
                                b1 - b0, n, begins, lrs, wds, float(self.param_groups[0]["momentum"]), 1.0,
                                self.lr_scale.data_ptr(), st), "tf_sgd_step")

    @torch.no_grad()
    def step(self, closure=None):
        for k in range(len(self.flat.buckets)):
            self.step_bucket(k)

    def steplr(self, step_size, gamma, advance=1):
        """scheduler.step() of main.py:81-83 on the device: epoch += advance; lr_scale = gamma ** (epoch // step_size)."""
        check(lib().tf_steplr_update(self.lr_scale.data_ptr(), self.epoch.data_ptr(), int(step_size), float(gamma), int(advance),
                                     stream_ptr(self.flat.device)), "tf_steplr_update")


def reduce_and_step(optimizer, group=None):
    """The bucket pipeline after a flat backward: for every bucket, on the communication stream,
    wait(bucket event) -> SUM all-reduce (world > 1) -> SGD.  The caller's stream joins at the end."""
    f = optimizer.flat
    cur = torch.cuda.current_stream(f.device)
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    with torch.cuda.stream(f.comm_stream):
        for k, (b0, b1, _blk) in enumerate(f.buckets):
            f.comm_stream.wait_event(f.events[k])
            if world > 1:
                dist.all_reduce(f.flat_grad[b0:b1], op=dist.ReduceOp.SUM, group=group)
            optimizer.step_bucket(k, f.comm_stream)
    cur.wait_stream(f.comm_stream)
