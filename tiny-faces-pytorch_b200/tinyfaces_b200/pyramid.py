"""GPU image pyramid level with the reference's exact CPU arithmetic (evaluation.py:40-50):
to_pil_image -> PIL bilinear resize -> ToTensor -> Normalize, bit-identical, without leaving the device.

The resampling coefficient tables are computed here on the host in float64 exactly as Pillow's
``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (bilinear filter, support 1, 22-bit fixed point); the kernels in
csrc/tf_pyramid.cu then reproduce Pillow's two integer passes (horizontal, uint8 intermediate, vertical).
"""
import ctypes
import math

import numpy as np
import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr

PRECISION_BITS = 32 - 8 - 2


def resized_size(w, h, size):
    """torchvision.transforms.functional.resize(img, int): the shorter side becomes `size`."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long_ / short)
    return (new_short, new_long) if w <= h else (new_long, new_short)


def bilinear_coeffs(in_size, out_size):
    """Pillow precompute_coeffs(inSize, 0, inSize, outSize, BILINEAR) + normalize_coeffs_8bpc -> (bounds, kk, ksize).
    Vectorised over the output index; the tap loop stays sequential so every float64 operation (and the order of
    the running weight sum) is the one Pillow's C code performs."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xx = np.arange(out_size, dtype=np.float64)
    center = 0.0 + (xx + 0.5) * scale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)          # (int) cast: truncation toward zero
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    k = np.zeros((out_size, ksize), np.float64)
    ww = np.zeros(out_size, np.float64)
    for x in range(ksize):
        a = np.abs((x + xmin - center + 0.5) * ss)
        w = np.where(a < 1.0, 1.0 - a, 0.0)
        w = np.where(x < xmax, w, 0.0)
        k[:, x] = w
        ww = np.where(x < xmax, ww + w, ww)
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz, None]
    kk = np.where(k < 0, np.trunc(-0.5 + k * (1 << PRECISION_BITS)), np.trunc(0.5 + k * (1 << PRECISION_BITS))).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kk, ksize


_TABLES = {}


def _tables(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    t = _TABLES.get(key)
    if t is None:
        b, k, ks = bilinear_coeffs(in_size, out_size)
        t = (torch.from_numpy(b).to(device), torch.from_numpy(k).to(device), ks)
        _TABLES[key] = t
    return t


@ops._on_tensor_device
def pyramid_level(img, size, mean, std):
    """img: CUDA float32 [3,H,W] in [0,1]; returns the normalised [1,3,H',W'] level whose shorter side is `size`."""
    _, H, W = img.shape
    Wo, Ho = resized_size(W, H, size)
    dev = img.device
    out = torch.empty((1, 3, Ho, Wo), dtype=torch.float32, device=dev)
    bh = kh = bv = kv = None
    ksh = ksv = 0
    if (Wo, Ho) != (W, H):              # PIL resamples an axis only when its size changes
        if Wo != W:
            bh, kh, ksh = _tables(W, Wo, dev)
        if Ho != H:
            bv, kv, ksv = _tables(H, Ho, dev)
    sz = ctypes.c_size_t()
    check(lib().tf_pyramid_workspace_bytes(H, W, Wo, ctypes.byref(sz)), "tf_pyramid_workspace_bytes")
    ws = ops._workspace(dev, sz.value)
    m = (ctypes.c_float * 3)(*[float(np.float32(v)) for v in mean])
    s = (ctypes.c_float * 3)(*[float(np.float32(v)) for v in std])
    check(lib().tf_pyramid_level(ptr(img.contiguous()), H, W, Ho, Wo, ptr(bh), ptr(kh), ksh, ptr(bv), ptr(kv), ksv, m, s,
                                 ptr(out), ptr(ws), ws.numel(), stream_ptr(dev)), "tf_pyramid_level")
    return out
