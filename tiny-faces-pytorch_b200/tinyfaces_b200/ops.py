"""Thin torch-tensor wrappers over the C-ABI entry points (one function per entry point).

PyTorch only provides device memory and the current stream; every kernel is in libtinyfaces_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr

_WS = {}


def _workspace(device, nbytes):
    """Grow-only per-device scratch buffer (caller-owned memory convention of the C-ABI)."""
    key = (device.type, device.index)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


# ----------------------------------------------------------------------------- NMS
def nms_device(boxes, scores, iou_threshold):
    """boxes [N,4], scores [N] CUDA float64/float32 -> (keep int64 [N] buffer, count int64 [1]) on device."""
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
    ws = _workspace(boxes.device, sz.value)
    check(lib().tf_nms(ptr(boxes), ptr(scores), n, eb, float(iou_threshold), ptr(keep), ptr(count), ptr(ws),
                       ws.numel(), stream_ptr(boxes.device)), "tf_nms")
    return keep, count


# ----------------------------------------------------------------------------- decode
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


def detloss_sample_device_(labels, max_pos, max_neg, seed):
    """In-place device balance sampler on labels [B,T,H,W]."""
    require_cuda(labels, "labels")
    B = labels.shape[0]
    L = labels[0].numel()
    sz = ctypes.c_size_t()
    check(lib().tf_detloss_sample_workspace_bytes(B, ctypes.byref(sz)), "tf_detloss_sample_workspace_bytes")
    ws = _workspace(labels.device, sz.value)
    check(lib().tf_detloss_sample_device(ptr(labels), B, L, int(max_pos), int(max_neg), int(seed) & (2**64 - 1),
                                         ptr(ws), ws.numel(), stream_ptr(labels.device)), "tf_detloss_sample_device")
    return labels


# ----------------------------------------------------------------------------- convolution GEMMs
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


def conv2d_wgrad_nhwc_strided(x, dy, ksize, stride):
    require_cuda(x, "x")
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    dw = torch.zeros((Cout, ksize * ksize, Cin), dtype=torch.float32, device=x.device)
    check(lib().tf_conv2d_wgrad_nhwc_strided(ptr(x), ptr(dy), B, H, W, Cin, Cout, ksize, stride, ptr(dw),
                                             stream_ptr(x.device)), "tf_conv2d_wgrad_nhwc_strided")
    return dw


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
