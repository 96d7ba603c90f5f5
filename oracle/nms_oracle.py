"""Loader for the plain-C NMS oracle + a pure-numpy version for small cases.
TEST INFRASTRUCTURE ONLY.  Semantics: see nms_oracle.c header."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libtf_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = ctypes.CDLL(so)
        _LIB.tf_oracle_nms_f64.restype = ctypes.c_int64
        _LIB.tf_oracle_nms_f64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_double, ctypes.c_void_p]
    return _LIB


def nms(boxes, scores, thr):
    """boxes [N,4] float64, scores [N] float64 -> int64 keep indices (C oracle)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float64)
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = _lib().tf_oracle_nms_f64(boxes.ctypes.data, scores.ctypes.data, n, float(thr), keep.ctypes.data)
    return keep[:k].copy()


def nms_numpy(boxes, scores, thr):
    """Pure-numpy greedy NMS (small N only)."""
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    scores = np.asarray(scores, dtype=np.float64)
    order = np.argsort(-scores, kind="stable")
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    dead = np.zeros(len(scores), bool)
    keep = []
    for a, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        rest = order[a + 1:]
        w = np.maximum(np.minimum(boxes[i, 2], boxes[rest, 2]) - np.maximum(boxes[i, 0], boxes[rest, 0]), 0)
        h = np.maximum(np.minimum(boxes[i, 3], boxes[rest, 3]) - np.maximum(boxes[i, 1], boxes[rest, 1]), 0)
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            ovr = inter / (area[i] + area[rest] - inter)
        dead[rest[ovr > thr]] = True
    return np.asarray(keep, dtype=np.int64)
