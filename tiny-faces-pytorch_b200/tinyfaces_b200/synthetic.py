"""Seeded synthetic inputs for benchmarks and examples (no dataset / checkpoint exists offline).
Shapes and distributions follow BASELINE.md section 2: images ~ randn, class_map ~ {-1: .98, 0: .01, +1: .01},
regression_map ~ 0.2 * randn, NMS boxes: centres U(0,S)^2, sizes U(10,70)^2, scores U(0,1) with duplicates."""
import math

import numpy as np
import torch


def targets(B, H3, W3, T=25, seed=0, p_neg=0.98, p_pos=0.01):
    r = np.random.RandomState(seed)
    u = r.rand(B, T, H3, W3)
    cm = np.zeros((B, T, H3, W3), np.float32)
    cm[u < p_neg] = -1
    cm[u > 1 - p_pos] = 1
    rm = (0.2 * r.randn(B, 4 * T, H3, W3)).astype(np.float32)
    return torch.from_numpy(cm), torch.from_numpy(rm)


def images(B, H, W, seed=0):
    return torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed))


def boxes(n, seed=0, extent=None, dup_frac=0.01):
    r = np.random.RandomState(seed)
    if extent is None:
        extent = 40.0 * math.sqrt(n / 4.0)
    c = r.rand(n, 2) * extent
    wh = 10 + 60 * r.rand(n, 2)
    b = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
    s = r.rand(n)
    nd = int(n * dup_frac)
    if nd:
        src, dst = r.randint(0, n, nd), r.randint(0, n, nd)
        b[dst] = b[src]
        s[dst] = s[src]
    return torch.from_numpy(b.astype(np.float64)), torch.from_numpy(s.astype(np.float64))
