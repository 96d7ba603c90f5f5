"""Thin torch-tensor wrappers over the C-ABI entry points (one function per entry point).

PyTorch only provides device memory and the current stream; every kernel is in libtinyfaces_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr

_WS = {}


def _on_tensor_device(fn):
    """Run the wrapped entry point with the CUDA device of its first CUDA tensor argument current: the library creates
    its streams, tensor maps and per-device state on the CURRENT device, which need not be the tensors' device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _workspace(device, nbytes):
    """Grow-only per-device scratch buffer (caller-owned memory convention of the C-ABI)."""
    key = (device.type, device.index)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


# ----------------------------------------------------------------------------- NMS
@_on_tensor_device
def nms_device(boxes, scores, iou_threshold, algorithm=None, exact_workspace=False):
    """boxes [N,4], scores [N] CUDA float64/float32 -> (keep int64 [N] buffer, count int64 [1]) on device.
    Only enqueues work (tf_nms is stream-ordered).  count == -1: the sort-and-sweep path's conflict list did not fit the
    workspace -- call again with algorithm=1 (``nms_keep`` does that).  algorithm: None = the library's automatic choice."""
    require_cuda(boxes, "boxes")
    require_cuda(scores, "scores")
    if boxes.dtype != scores.dtype:
        raise RuntimeError("nms: boxes and scores must have the same dtype")   # torchvision raises too
    if boxes.dtype not in (torch.float64, torch.float32):
        raise RuntimeError("nms: dtype must be float64 or float32")
    boxes = boxes.contiguous()
    scores = scores.contiguous()
    n = boxes.shape[0]
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=boxes.device)
    count = torch.zeros(1, dtype=torch.int64, device=boxes.device)
    if n == 0:
        return keep[:0], count
    eb = boxes.element_size()
    sz = ctypes.c_size_t()
    check(lib().tf_nms_workspace_bytes(n, eb, ctypes.byref(sz)), "tf_nms_workspace_bytes")
    # exact_workspace: a private buffer of exactly the reported size (the shared grow-only buffer may be larger, and the
    # sort-and-sweep path turns every extra byte into conflict-list capacity)
    ws = torch.empty(sz.value, dtype=torch.uint8, device=boxes.device) if exact_workspace else _workspace(boxes.device, sz.value)
    if algorithm is None:
        check(lib().tf_nms(ptr(boxes), ptr(scores), n, eb, float(iou_threshold), ptr(keep), ptr(count), ptr(ws),
                           ws.numel(), stream_ptr(boxes.device)), "tf_nms")
    else:
        check(lib().tf_nms_algo(ptr(boxes), ptr(scores), n, eb, float(iou_threshold), int(algorithm), ptr(keep), ptr(count),
                                ptr(ws), ws.numel(), stream_ptr(boxes.device)), "tf_nms_algo")
    return keep, count


def nms_keep(boxes, scores, iou_threshold, algorithm=None):
    """The kept indices as a device tensor [K] (reads the count: ONE host synchronisation).  Falls back to the blocked
    bit-matrix algorithm when the sort-and-sweep conflict list overflowed (count == -1)."""
    keep, count = nms_device(boxes, scores, iou_threshold, algorithm)
    k = int(count.item())
    if k < 0:
        keep, count = nms_device(boxes, scores, iou_threshold, 1)
        k = int(count.item())
    return keep[:k]


def nms_sweep_stats(n, elem_bytes, device):
    """{edges, pair_tests, rounds, edge_capacity} of the last sort-and-sweep tf_nms on this device's workspace (synchronises)."""
    ws = _WS[(device.type, device.index)]
    out = (ctypes.c_int64 * 4)()
    check(lib().tf_nms_sweep_stats(n, elem_bytes, ptr(ws), ws.numel(), out, stream_ptr(device)), "tf_nms_sweep_stats")
    return dict(edges=out[0], pair_tests=out[1], rounds=out[2], edge_capacity=out[3])


# ----------------------------------------------------------------------------- decode
@_on_tensor_device
def decode_device(cls, reg, prob, cls_strides, reg_strides, B, H, W, T, templates, prob_thresh, invalid_x_mask,
                  invalid_t_mask, rf, scale, capacity=None, want_src=False):
    """Returns (boxes f64 [cap,4], scores f64 [cap], src int64 [cap] | None, count int64 [1]) on device.
    `cls`, `reg`, `prob` are CUDA float32 tensors addressed through explicit (b, y, x, c) element strides."""
    require_cuda(cls, "cls")
    dev = cls.device
    if capacity is None:
        capacity = B * H * W * T
    boxes = torch.empty((max(capacity, 1), 4), dtype=torch.float64, device=dev)
    scores = torch.empty(max(capacity, 1), dtype=torch.float64, device=dev)
    src = torch.empty(max(capacity, 1), dtype=torch.int64, device=dev) if want_src else None
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    sz = ctypes.c_size_t()
    check(lib().tf_decode_workspace_bytes(B, H, W, ctypes.byref(sz)), "tf_decode_workspace_bytes")
    ws = _workspace(dev, sz.value)
    tpl = np.ascontiguousarray(templates, dtype=np.float64)
    cs = (ctypes.c_int64 * 4)(*[int(v) for v in cls_strides])
    rs = (ctypes.c_int64 * 4)(*[int(v) for v in reg_strides])
    st = (ctypes.c_int64 * 2)(int(rf["stride"][0]), int(rf["stride"][1]))
    of = (ctypes.c_int64 * 2)(int(rf["offset"][0]), int(rf["offset"][1]))
    check(lib().tf_decode(ptr(cls), ptr(reg), ptr(prob), cs, rs, B, H, W, T,
                          tpl.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), float(np.float32(prob_thresh)),
                          int(invalid_x_mask), int(invalid_t_mask), st, of, float(scale), ptr(boxes), ptr(scores),
                          ptr(src), capacity, ptr(count), ptr(ws), ws.numel(), stream_ptr(dev)), "tf_decode")
    return boxes, scores, src, count


# ----------------------------------------------------------------------------- loss
@_on_tensor_device
def detloss_ohem_(output, class_map, thresh=0.03):
    """In-place OHEM on class_map (loss.py:59-63)."""
    require_cuda(output, "output")
    require_cuda(class_map, "class_map")
    B, C, H, W = output.shape
    T = class_map.shape[1]
    assert output.is_contiguous() and class_map.is_contiguous() and output.dtype == torch.float32
    check(lib().tf_detloss_ohem(ptr(output), ptr(class_map), B, T, H * W, float(thresh), stream_ptr(output.device)),
          "tf_detloss_ohem")
    return class_map


@_on_tensor_device
def detloss_fwd_bwd(output, labels, regression_map, reg_weight=1.0):
    """Returns (sums float64 [2] = (cls_sum, reg_sum), grad float32 like output)."""
    require_cuda(output, "output")
    B, C, H, W = output.shape
    T = labels.shape[1]
    assert output.is_contiguous() and labels.is_contiguous() and regression_map.is_contiguous()
    grad = torch.empty_like(output)
    sums = torch.zeros(2, dtype=torch.float64, device=output.device)
    check(lib().tf_detloss_fwd_bwd(ptr(output), ptr(labels), ptr(regression_map), B, T, H * W, float(reg_weight),
                                   ptr(grad), ptr(sums), stream_ptr(output.device)), "tf_detloss_fwd_bwd")
    return sums, grad


@_on_tensor_device
def detloss_sample_device_(labels, max_pos, max_neg, seed, counter=None):
    """In-place device balance sampler on labels [B,T,H,W].  counter: optional device int64 [1] draw counter (mixed into
    the seed, incremented by the call)."""
    require_cuda(labels, "labels")
    B = labels.shape[0]
    L = labels[0].numel()
    sz = ctypes.c_size_t()
    check(lib().tf_detloss_sample_workspace_bytes(B, ctypes.byref(sz)), "tf_detloss_sample_workspace_bytes")
    ws = _workspace(labels.device, sz.value)
    check(lib().tf_detloss_sample_device_ctr(ptr(labels), B, L, int(max_pos), int(max_neg), int(seed) & (2**64 - 1), ptr(counter),
                                             ptr(ws), ws.numel(), stream_ptr(labels.device)), "tf_detloss_sample_device")
    return labels


# ----------------------------------------------------------------------------- convolution GEMMs
@_on_tensor_device
def conv2d_nhwc(x, w_packed, ksize, bias=None, x_lo=None, w_lo=None, out=None):
    """x [B,H,W,Cin] fp32 NHWC, w_packed [Cout, k*k, Cin] -> y [B,H,W,Cout] (stride 1, same padding)."""
    require_cuda(x, "x")
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    assert x.is_contiguous() and w_packed.is_contiguous() and w_packed.shape[1] == ksize * ksize and w_packed.shape[2] == Cin
    y = out if out is not None else torch.empty((B, H, W, Cout), dtype=torch.float32, device=x.device)
    check(lib().tf_conv2d_nhwc(ptr(x), ptr(x_lo), B, H, W, Cin, ptr(w_packed), ptr(w_lo), Cout, ksize, ptr(bias),
                               ptr(y), stream_ptr(x.device)), "tf_conv2d_nhwc")
    return y


@_on_tensor_device
def conv2d_wgrad_nhwc(x, dy, ksize, out=None):
    """dw_packed [Cout, k*k, Cin] = sum_pixels dy (x) x (stride 1, same padding)."""
    require_cuda(x, "x")
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    assert x.is_contiguous() and dy.is_contiguous()
    dw = out if out is not None else torch.zeros((Cout, ksize * ksize, Cin), dtype=torch.float32, device=x.device)
    check(lib().tf_conv2d_wgrad_nhwc(ptr(x), ptr(dy), B, H, W, Cin, Cout, ksize, ptr(dw), stream_ptr(x.device)),
          "tf_conv2d_wgrad_nhwc")
    return dw


@_on_tensor_device
def conv2d_nhwc_strided(x, w_packed, ksize, stride, bias=None):
    """Strided variant (TMA traversal stride): y [B, ceil(H/s), ceil(W/s), Cout]."""
    require_cuda(x, "x")
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    check(lib().tf_conv2d_nhwc_strided(ptr(x), B, H, W, Cin, ptr(w_packed), Cout, ksize, stride, ptr(bias), ptr(y),
                                       stream_ptr(x.device)), "tf_conv2d_nhwc_strided")
    return y


@_on_tensor_device
def conv2d_wgrad_nhwc_strided(x, dy, ksize, stride):
    require_cuda(x, "x")
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    dw = torch.zeros((Cout, ksize * ksize, Cin), dtype=torch.float32, device=x.device)
    check(lib().tf_conv2d_wgrad_nhwc_strided(ptr(x), ptr(dy), B, H, W, Cin, Cout, ksize, stride, ptr(dw),
                                             stream_ptr(x.device)), "tf_conv2d_wgrad_nhwc_strided")
    return dw


@_on_tensor_device
def conv2d_dgrad_s2_nhwc(dy, w_packed, ksize, H, W, out=None, accumulate=False):
    """Input gradient of a stride-2 conv: dy [B, ceil(H/2), ceil(W/2), Cdy], w_packed [Cdx, k*k (flipped), Cdy] -> dx [B,H,W,Cdx]."""
    require_cuda(dy, "dy")
    B, Ho, Wo, Cdy = dy.shape
    Cdx = w_packed.shape[0]
    assert dy.is_contiguous() and w_packed.is_contiguous() and Ho == (H + 1) // 2 and Wo == (W + 1) // 2
    dx = out if out is not None else torch.zeros((B, H, W, Cdx), dtype=torch.float32, device=dy.device)
    check(lib().tf_conv2d_dgrad_s2_nhwc(ptr(dy), B, H, W, Cdy, ptr(w_packed), Cdx, ksize, int(bool(accumulate)), ptr(dx),
                                        stream_ptr(dy.device)), "tf_conv2d_dgrad_s2_nhwc")
    return dx


def gemm_error_flag():
    v = ctypes.c_int(0)
    check(lib().tf_gemm_error_flag(ctypes.byref(v)), "tf_gemm_error_flag")
    return v.value


# ----------------------------------------------------------------------------- elementwise kernels, one by one
def _bn_ws(device):
    sz = ctypes.c_size_t()
    check(lib().tf_bn_workspace_bytes(ctypes.byref(sz)), "tf_bn_workspace_bytes")
    return _workspace(device, sz.value)


@_on_tensor_device
def bn_train_fwd(y, gamma, beta, run_mean=None, run_var=None, res=None, relu=True, eps=1e-5, momentum=0.1, round_tf32=False,
                 want_mask=True):
    """y [..., C] NHWC fp32 -> (out, relu_mask uint32 | None, save_mean, save_rstd); running stats updated in place."""
    require_cuda(y, "y")
    C = y.shape[-1]
    M = y.numel() // C
    assert y.is_contiguous() and y.dtype == torch.float32
    out = torch.empty_like(y)
    mask = torch.zeros((M * C + 31) // 32, dtype=torch.int32, device=y.device) if want_mask else None
    mean = torch.empty(C, dtype=torch.float32, device=y.device)
    rstd = torch.empty(C, dtype=torch.float32, device=y.device)
    ws = _bn_ws(y.device)
    check(lib().tf_bn_train_fwd(ptr(y), M, C, ptr(gamma), ptr(beta), float(eps), float(momentum), ptr(run_mean), ptr(run_var),
                                ptr(res), int(relu), int(round_tf32), ptr(out), ptr(mask), ptr(mean), ptr(rstd), ptr(ws),
                                ws.numel(), stream_ptr(y.device)), "tf_bn_train_fwd")
    return out, mask, mean, rstd


@_on_tensor_device
def bn_eval_fwd(y, gamma, beta, run_mean, run_var, res=None, relu=True, eps=1e-5, round_tf32=False):
    require_cuda(y, "y")
    C = y.shape[-1]
    M = y.numel() // C
    out = torch.empty_like(y)
    ws = _bn_ws(y.device)
    check(lib().tf_bn_eval_fwd(ptr(y), M, C, ptr(gamma), ptr(beta), ptr(run_mean), ptr(run_var), float(eps), ptr(res), int(relu),
                               int(round_tf32), ptr(out), ptr(ws), ws.numel(), stream_ptr(y.device)), "tf_bn_eval_fwd")
    return out


@_on_tensor_device
def bn_bwd(dout, relu_mask, y, save_mean, save_rstd, gamma, want_g=False, round_tf32=False):
    """-> (dy, dgamma, dbeta, g | None)."""
    require_cuda(dout, "dout")
    C = y.shape[-1]
    M = y.numel() // C
    dy = torch.empty_like(y)
    dgamma = torch.empty(C, dtype=torch.float32, device=y.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=y.device)
    g = torch.empty_like(y) if want_g else None
    ws = _bn_ws(y.device)
    check(lib().tf_bn_bwd(ptr(dout), ptr(relu_mask), ptr(y), ptr(save_mean), ptr(save_rstd), ptr(gamma), M, C, ptr(dgamma),
                          ptr(dbeta), ptr(dy), ptr(g), int(round_tf32), ptr(ws), ws.numel(), stream_ptr(y.device)), "tf_bn_bwd")
    return dy, dgamma, dbeta, g


@_on_tensor_device
def maxpool_fwd(x):
    """x [B,H,W,C] NHWC -> (out [B,Ho,Wo,C], argmax uint8)."""
    require_cuda(x, "x")
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty((B, Ho, Wo, C), dtype=torch.float32, device=x.device)
    am = torch.empty((B, Ho, Wo, C), dtype=torch.uint8, device=x.device)
    check(lib().tf_maxpool_fwd(ptr(x), B, H, W, C, ptr(out), ptr(am), stream_ptr(x.device)), "tf_maxpool_fwd")
    return out, am


@_on_tensor_device
def maxpool_bwd(argmax, dout, H, W):
    require_cuda(dout, "dout")
    B, Ho, Wo, C = dout.shape
    dx = torch.empty((B, H, W, C), dtype=torch.float32, device=dout.device)
    check(lib().tf_maxpool_bwd(ptr(argmax), ptr(dout), B, H, W, C, ptr(dx), stream_ptr(dout.device)), "tf_maxpool_bwd")
    return dx


def _head_ws(device, Cn):
    sz = ctypes.c_size_t()
    check(lib().tf_head_workspace_bytes(Cn, ctypes.byref(sz)), "tf_head_workspace_bytes")
    return _workspace(device, sz.value)


@_on_tensor_device
def head_upsample_add_fwd(s3, s4, up_w, Cn):
    """s3 [B,H3,W3,Cp], s4 [B,H4,W4,Cp] NHWC, up_w [Cn,Cn,4,4] -> out [B,Cn,H3,W3] NCHW."""
    require_cuda(s3, "s3")
    B, H3, W3, Cp = s3.shape
    H4, W4 = s4.shape[1], s4.shape[2]
    out = torch.empty((B, Cn, H3, W3), dtype=torch.float32, device=s3.device)
    ws = _head_ws(s3.device, Cn)
    check(lib().tf_head_upsample_add_fwd(ptr(s3), ptr(s4), ptr(up_w), B, H3, W3, H4, W4, Cn, Cp, ptr(out), ptr(ws), ws.numel(),
                                         stream_ptr(s3.device)), "tf_head_upsample_add_fwd")
    return out


@_on_tensor_device
def head_upsample_add_bwd(dout, up_w, H4, W4, Cp):
    require_cuda(dout, "dout")
    B, Cn, H3, W3 = dout.shape
    ds3 = torch.empty((B, H3, W3, Cp), dtype=torch.float32, device=dout.device)
    ds4 = torch.empty((B, H4, W4, Cp), dtype=torch.float32, device=dout.device)
    ws = _head_ws(dout.device, Cn)
    check(lib().tf_head_upsample_add_bwd(ptr(dout), ptr(up_w), B, H3, W3, H4, W4, Cn, Cp, ptr(ds3), ptr(ds4), ptr(ws), ws.numel(),
                                         stream_ptr(dout.device)), "tf_head_upsample_add_bwd")
    return ds3, ds4


@_on_tensor_device
def conv2d_nhwc_res(x, w_packed, res, res_mask=None):
    """1x1: y = x @ w^T + (mask bit ? res : 0); x [B,H,W,Cin], w_packed [Cout,1,Cin], res [B,H,W,Cout]."""
    require_cuda(x, "x")
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    y = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x.device)
    check(lib().tf_conv2d_nhwc_res(ptr(x), B, H, W, Cin, ptr(w_packed), Cout, ptr(res), ptr(res_mask), ptr(y),
                                   stream_ptr(x.device)), "tf_conv2d_nhwc_res")
    return y
