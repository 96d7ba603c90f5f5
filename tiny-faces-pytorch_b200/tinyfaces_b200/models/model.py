"""DetectionModel -- drop-in for /root/reference/tinyfaces/models/model.py:7-128.

Same constructor, same parameter / buffer names (``state_dict`` is key-compatible, including the unused
``model.fc`` the reference keeps), same ``learnable_parameters`` groups, same ``forward(x) -> [B, 5T, H/8, W/8]``
NCHW fp32 contract, and it participates in autograd.  The arithmetic, however, never touches torch.nn: the
whole trunk + heads forward and backward run inside libtinyfaces_b200.so (tcgen05 implicit-GEMM convolutions,
fused BN / ReLU / residual kernels) through ``tf_model_forward`` / ``tf_model_backward``.  There is no CPU path.

Restrictions that differ from an ordinary nn.Module (both raise instead of computing something wrong):
  * one grad-enabled forward in flight per model: the backward's saved activations live in the executor's single
    workspace, so ``backward()`` must run before the next ``forward()`` of the same model;
  * ``precision``: "fast" (default) = one TF32 tensor-core product per GEMM -- what the reference itself runs on CUDA
    (torch.backends.cudnn.allow_tf32 defaults to True); measured against the fp32 CPU reference on the conditioned
    synthetic weights its score map is off by ~5e-3 (max-norm) and its weight gradients by 20-30 % rel-L2 (error
    amplification of the untrained net).  "parity" = 3xTF32 (fp32-equivalent products): score map within
    1e-3, ~2x the step time, 2x the activation memory.  "mixed" = the 3xTF32 forward of "parity" (same score map, exact
    ReLU masks and batch statistics) with the single-product TF32 backward of "fast" (what cuDNN's TF32 default does to
    the reference's own backward on CUDA); inference-only use is identical to "parity".
"""
import ctypes

import numpy as np
import torch
from torch import nn
from torchvision.models import ResNet101_Weights, resnet101

from .. import _lib
from .._lib import check, lib, stream_ptr

PRECISION = {"fast": 1, "parity": 2, "mixed": 3}


class _Executor:
    """Owns the C-side model handle, the workspace and the parameter pointer table."""

    def __init__(self, num_templates):
        h = ctypes.c_void_p()
        check(lib().tf_model_create(num_templates, ctypes.byref(h)), "tf_model_create")
        self.handle = h
        n = lib().tf_model_num_params(h)
        self.names = [lib().tf_model_param_name(h, i).decode() for i in range(n)]
        self.workspace = None
        self._sizes = {}
        self._out_shapes = {}
        # Every training forward overwrites the activations / BN statistics / ReLU masks the backward reads (they live in
        # the ONE workspace).  `generation` counts forwards; an autograd node remembers the generation it belongs to and
        # refuses to back-propagate through a workspace that a later forward has overwritten.
        self.generation = 0

    def __del__(self):
        try:
            lib().tf_model_destroy(self.handle)
        except Exception:
            pass

    def ensure_workspace(self, device, B, H, W, training, mode):
        key = (B, H, W, bool(training), mode)
        need = self._sizes.get(key)
        if need is None:
            sz = ctypes.c_size_t()
            check(lib().tf_model_workspace_bytes(self.handle, B, H, W, int(training), mode, ctypes.byref(sz)),
                  "tf_model_workspace_bytes")
            need = self._sizes[key] = sz.value
        if self.workspace is None or self.workspace.numel() < need or self.workspace.device != device:
            self.workspace = None
            self.workspace = torch.empty(need + 4096, dtype=torch.uint8, device=device)
        return self.workspace


def _run_forward(module, x):
    """tf_model_forward on the module's current parameters (no autograd bookkeeping)."""
    ex = module._executor
    _lib.require_cuda(x, "x")
    if x.dtype != torch.float32:
        raise RuntimeError("DetectionModel expects float32 input")
    x = x.contiguous()
    B, _, H, W = x.shape
    mode = PRECISION[module.precision]
    training = bool(module.training)
    ws = ex.ensure_workspace(x.device, B, H, W, training, mode)
    table = module._tables()[0]
    cached = module.__dict__.get("_ptr_table")
    if cached is None or cached[0] is not table or cached[1] != x.device:
        for t in table:
            if t.device != x.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("DetectionModel parameters must be contiguous float32 tensors on the input's device")
        # the ctypes pointer table of the ~570 parameters / buffers costs ~0.3 ms to build: once per parameter set
        # (optimizers update parameters in place, so the addresses are stable; _apply() drops the cache)
        flat = module.__dict__.get("_flat")
        if flat is not None and not flat.is_valid():
            raise RuntimeError("DetectionModel: the parameters were moved / re-allocated after the flat parameter store "
                               "(tinyfaces_b200.optim.FlatParams / FlatSGD) was built -- build the optimizer after model.to(device)")
        ptrs = (ctypes.c_void_p * len(table))(*[t.data_ptr() for t in table])
        cached = module.__dict__["_ptr_table"] = (table, x.device, ptrs)
        module.__dict__.pop("_graphs", None)             # captured inference graphs hold the old parameter addresses
    ptrs = cached[2]
    key = (H, W)
    shp = ex._out_shapes.get(key)
    if shp is None:
        h3, w3 = ctypes.c_int(), ctypes.c_int()
        check(lib().tf_model_output_shape(ex.handle, H, W, ctypes.byref(h3), ctypes.byref(w3)), "tf_model_output_shape")
        shp = ex._out_shapes[key] = (h3.value, w3.value)
    out = torch.empty((B, 5 * module.num_templates, shp[0], shp[1]), dtype=torch.float32, device=x.device)
    ex.generation += 1
    with torch.cuda.device(x.device):       # the library creates streams / tensor maps on the CURRENT device
        check(lib().tf_model_forward(ex.handle, x.data_ptr(), B, H, W, ptrs, int(training), mode, float(module.bn_momentum),
                                     out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(x.device)), "tf_model_forward")
    if training:
        torch._foreach_add_(module._bn_counters(), 1)        # num_batches_tracked of all 94 BN layers, one launch
    return out


class _TrunkFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, *tensors):
        out = _run_forward(module, x)
        ctx.module = module
        ctx.generation = module._executor.generation
        return out

    @staticmethod
    def backward(ctx, grad_out):
        module = ctx.module
        ex = module._executor
        if ctx.generation != ex.generation:
            raise RuntimeError(
                "DetectionModel.backward: the forward this graph node belongs to (#%d) is no longer the executor's last "
                "forward (#%d).  The activations saved for the backward live in ONE workspace that every forward overwrites: "
                "only one forward may be in flight per model -- call backward() before the next forward() (or use a second "
                "DetectionModel instance sharing the parameters)." % (ctx.generation, ex.generation))
        grad_out = grad_out.contiguous()
        flat = module.__dict__.get("_flat")
        if flat is not None:
            # flat-gradient fast path (tinyfaces_b200.optim.FlatParams): the library writes every gradient straight into the
            # persistent flat buffer the .grad tensors view (OVERWRITING: zero_grad() + backward() semantics) and records the
            # bucket events the overlapped all-reduce / SGD pipeline waits on
            table = module.__dict__.get("_flat_grad_table")
            if table is None or table[0] is not flat:
                table = module.__dict__["_flat_grad_table"] = (flat, flat.grad_pointer_table(ex.names))
            with torch.cuda.device(grad_out.device):
                check(lib().tf_model_backward_ex(ex.handle, grad_out.data_ptr(), table[1], len(flat.events), flat._ev_ptrs,
                                                 flat._ev_blocks, stream_ptr(grad_out.device)), "tf_model_backward_ex")
            for p, g in zip(flat.params, flat.grad_views):
                if p.grad is not g:
                    p.grad = g
            return (None, None) + (None,) * len(module._autograd_names)
        named = dict(zip(module._tables()[2], module._tables()[1]))
        grads = {}
        ptrs = (ctypes.c_void_p * len(ex.names))()
        for i, name in enumerate(ex.names):
            p = named.get(name)
            if p is not None and p.requires_grad and name != "score4_upsample.weight":
                g = torch.empty_like(p)
                grads[name] = g
                ptrs[i] = g.data_ptr()
            else:
                ptrs[i] = None
        with torch.cuda.device(grad_out.device):
            check(lib().tf_model_backward(ex.handle, grad_out.data_ptr(), ptrs, stream_ptr(grad_out.device)),
                  "tf_model_backward")
        out = [None, None]
        for name in module._autograd_names:
            out.append(grads.get(name))
        return tuple(out)


class DetectionModel(nn.Module):
    """Hybrid-resolution Tiny Faces detector (model.py:7-40), Blackwell-native execution."""

    def __init__(self, base_model=resnet101, pretrained_weights=ResNet101_Weights.IMAGENET1K_V1, num_templates=1,
                 num_objects=1):
        super().__init__()
        if num_objects != 1:
            raise RuntimeError("tinyfaces_b200 supports num_objects == 1 (the reference's only use, main.py:52)")
        output = (num_objects + 4) * num_templates            # model.py:19
        self.num_templates = num_templates
        self.model = base_model(weights=pretrained_weights)   # parameter container only -- never called
        del self.model.layer4                                 # model.py:23
        self.score_res3 = nn.Conv2d(512, output, kernel_size=1, padding=0)
        self.score_res4 = nn.Conv2d(1024, output, kernel_size=1, padding=0)
        self.score4_upsample = nn.ConvTranspose2d(output, output, kernel_size=4, stride=2, padding=1, bias=False)
        self._init_bilinear()
        self.precision = "fast"
        self.bn_momentum = 0.1                                # nn.BatchNorm2d default
        self.cuda_graphs = False                              # eval / no_grad forwards of small inputs replay a captured CUDA graph
        self.cuda_graph_max_pixels = 1300 * 1300              # (larger levels are not launch-bound)
        self._executor_obj = None
        self._checked_upsample = False

    @property
    def _executor(self):
        if self._executor_obj is None:
            object.__setattr__(self, "_executor_obj", _Executor(self.num_templates))
        return self._executor_obj

    def _init_bilinear(self):
        """model.py:45-65: diagonal bilinear 2x kernel with 1-D taps [.25, .75, .75, .25]."""
        k = self.score4_upsample.kernel_size[0]
        factor = np.floor((k + 1) / 2)
        center = factor if k % 2 == 1 else factor + 0.5
        taps = 1.0 - np.abs(np.arange(1, k + 1) - center) / factor
        w = torch.zeros_like(self.score4_upsample.weight)
        ch = torch.arange(w.shape[0])
        w[ch, ch] = torch.tensor(np.outer(taps, taps), dtype=w.dtype)
        self.score4_upsample.weight = nn.Parameter(w)

    def learnable_parameters(self, lr):
        """model.py:67-87 -- the four optimizer groups."""
        return [{'params': self.model.parameters(), 'lr': lr},
                {'params': self.score_res3.parameters(), 'lr': 0.1 * lr},
                {'params': self.score_res4.parameters(), 'lr': 1 * lr},
                {'params': self.score4_upsample.parameters(), 'lr': 0}]

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() may replace parameter storage: drop the cached tensor tables
        self.__dict__.pop("_cache", None)
        self.__dict__.pop("_bn_cnt", None)
        self.__dict__.pop("_ptr_table", None)
        self.__dict__.pop("_graphs", None)
        return super()._apply(fn, *args, **kwargs)

    def _bn_counters(self):
        c = self.__dict__.get("_bn_cnt")
        if c is None:
            c = [m.num_batches_tracked for m in self.modules()
                 if isinstance(m, nn.BatchNorm2d) and m.num_batches_tracked is not None]
            self.__dict__["_bn_cnt"] = c
        return c

    def _tables(self):
        """(parameter/buffer tensors in executor order, autograd parameter list) -- cached, the walk over
        named_parameters() costs more host time than a small pyramid level takes on the GPU."""
        c = self.__dict__.get("_cache")
        if c is None:
            pnamed = dict(self.named_parameters())
            names = self._executor.names
            params = [pnamed[n] for n in names if n in pnamed]
            named = dict(pnamed)
            named.update(dict(self.named_buffers()))
            c = ([named[n] for n in names], params, [n for n in names if n in pnamed])
            self.__dict__["_cache"] = c
        return c

    def _tensor_table(self):
        return [t.data for t in self._tables()[0]]

    @property
    def _autograd_names(self):
        return self._tables()[2]

    def debug_tensor(self, name):
        """Test hook: NHWC copy of an internal activation of the last forward (see tf_model_get_tensor)."""
        shape = (ctypes.c_int * 4)()
        ex = self._executor
        dev = ex.workspace.device
        check(lib().tf_model_get_tensor(ex.handle, name.encode(), None, 0, shape, stream_ptr(dev)), "tf_model_get_tensor")
        t = torch.empty(tuple(shape), dtype=torch.float32, device=dev)
        check(lib().tf_model_get_tensor(ex.handle, name.encode(), t.data_ptr(), t.numel(), shape, stream_ptr(dev)),
              "tf_model_get_tensor")
        return t

    def forward_train_flat(self, x):
        """Training forward WITHOUT autograd bookkeeping (pairs with backward_flat; needs a FlatParams / FlatSGD store)."""
        if self.__dict__.get("_flat") is None:
            raise RuntimeError("forward_train_flat needs tinyfaces_b200.optim.FlatParams / FlatSGD on this model")
        if not self.training:
            raise RuntimeError("forward_train_flat: the model is in eval mode")
        with torch.no_grad():
            return _run_forward(self, x)

    def backward_flat(self, grad_out):
        """Backward of the last forward_train_flat: every gradient is written into the flat store (overwriting) and the
        bucket events of the overlapped all-reduce / SGD pipeline are recorded.  No autograd engine involved."""
        flat = self.__dict__.get("_flat")
        ex = self._executor
        table = self.__dict__.get("_flat_grad_table")
        if table is None or table[0] is not flat:
            table = self.__dict__["_flat_grad_table"] = (flat, flat.grad_pointer_table(ex.names))
        grad_out = grad_out.contiguous()
        with torch.cuda.device(grad_out.device):
            check(lib().tf_model_backward_ex(ex.handle, grad_out.data_ptr(), table[1], len(flat.events), flat._ev_ptrs,
                                             flat._ev_blocks, stream_ptr(grad_out.device)), "tf_model_backward_ex")

    def _graph_forward(self, x):
        """Inference forward replayed from a CUDA graph captured per (input shape, precision, workspace): the small pyramid
        levels are launch-bound (~300 launches for < 1 ms of GPU work).  The returned tensor is the graph's static output:
        valid until the next forward of the same shape (get_detections consumes it, stream-ordered, before that)."""
        ex = self._executor
        graphs = self.__dict__.setdefault("_graphs", {})
        ws = ex.workspace
        key = (tuple(x.shape), self.precision, ws.data_ptr() if ws is not None else 0, x.device.index)
        g = graphs.get(key)
        if g is None:
            static_x = x.detach().clone().contiguous()
            _run_forward(self, static_x)                                  # eager: sizes the workspace, fills the library's caches
            ws = ex.workspace
            key = (tuple(x.shape), self.precision, ws.data_ptr(), x.device.index)
            for k in [k for k in graphs if k[2] != ws.data_ptr()]:        # graphs captured on a workspace that has been replaced
                del graphs[k]
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                static_out = _run_forward(self, static_x)
            g = graphs[key] = (graph, static_x, static_out)
        graph, static_x, static_out = g
        static_x.copy_(x, non_blocking=True)
        graph.replay()
        ex.generation += 1
        return static_out

    def forward(self, x):
        if torch.is_grad_enabled():
            out = _TrunkFunction.apply(x, self, *self._tables()[1])
        elif (self.cuda_graphs and not self.training and x.is_cuda and self._checked_upsample
              and x.shape[0] * x.shape[2] * x.shape[3] <= self.cuda_graph_max_pixels):
            out = self._graph_forward(x)
        else:
            out = _run_forward(self, x)      # torch.no_grad(): skip autograd.Function.apply over ~300 parameter tensors
        if not self._checked_upsample:
            v = ctypes.c_float()
            check(lib().tf_model_upsample_offdiag(self._executor.handle, ctypes.byref(v), stream_ptr(x.device)),
                  "tf_model_upsample_offdiag")
            if v.value != 0.0:
                raise RuntimeError("score4_upsample.weight is not diagonal (max off-diagonal %g): only the reference's "
                                   "frozen bilinear kernel (model.py:45-65) is supported" % v.value)
            self._checked_upsample = True
        return out
