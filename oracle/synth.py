"""Seeded synthetic weights / inputs shared by the golden generator, the parity
tests and bench.py (TEST INFRASTRUCTURE ONLY).

There is no dataset and no checkpoint offline, so every experiment uses a
synthetic ``state_dict`` with exactly the reference's 571 keys and shapes
(/root/reference/tinyfaces/models/model.py:12-40 + torchvision resnet101 minus
layer4) drawn from a torch CPU generator: convs ~ kaiming-normal(fan_out)
(torchvision/models/resnet.py:208-214), BN gamma=1 (bn3 gamma configurable --
SURVEY.md App. C), beta=0, heads ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)), the
upsample = model.py:45-65's diagonal bilinear kernel.
"""
import math

import numpy as np
import torch

from .model_oracle import LAYERS, bilinear_upsample_weight


def _conv(g, cout, cin, k):
    std = math.sqrt(2.0 / (cout * k * k))
    return torch.randn(cout, cin, k, k, generator=g) * std


def _bn(sd, prefix, c, gamma=1.0):
    sd[prefix + ".weight"] = torch.full((c,), float(gamma))
    sd[prefix + ".bias"] = torch.zeros(c)
    sd[prefix + ".running_mean"] = torch.zeros(c)
    sd[prefix + ".running_var"] = torch.ones(c)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def synthetic_state_dict(seed=0, num_templates=25, bn3_gamma=1.0, beta_jitter=0.0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    sd["model.conv1.weight"] = _conv(g, 64, 3, 7)
    _bn(sd, "model.bn1", 64)
    inplanes = 64
    for name, blocks, planes, _stride in LAYERS:
        for i in range(blocks):
            p = "model.%s.%d" % (name, i)
            sd[p + ".conv1.weight"] = _conv(g, planes, inplanes, 1)
            _bn(sd, p + ".bn1", planes)
            sd[p + ".conv2.weight"] = _conv(g, planes, planes, 3)
            _bn(sd, p + ".bn2", planes)
            sd[p + ".conv3.weight"] = _conv(g, planes * 4, planes, 1)
            _bn(sd, p + ".bn3", planes * 4, bn3_gamma)
            if i == 0:
                sd[p + ".downsample.0.weight"] = _conv(g, planes * 4, inplanes, 1)
                _bn(sd, p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    # unused-but-present torchvision leftovers (model.py:23 deletes only layer4)
    sd["model.fc.weight"] = (torch.rand(1000, 2048, generator=g) * 2 - 1) / math.sqrt(2048)
    sd["model.fc.bias"] = (torch.rand(1000, generator=g) * 2 - 1) / math.sqrt(2048)
    out = 5 * num_templates
    for nm, cin in (("score_res3", 512), ("score_res4", 1024)):
        b = 1.0 / math.sqrt(cin)
        sd[nm + ".weight"] = (torch.rand(out, cin, 1, 1, generator=g) * 2 - 1) * b
        sd[nm + ".bias"] = (torch.rand(out, generator=g) * 2 - 1) * b
    sd["score4_upsample.weight"] = bilinear_upsample_weight(out)
    if beta_jitter:
        for k in list(sd):
            if k.endswith(".bias") and ("bn" in k or "downsample.1" in k):
                sd[k] = torch.randn(sd[k].shape, generator=g) * beta_jitter
            if k.endswith(".weight") and sd[k].dim() == 1:
                sd[k] = sd[k] * (1 + beta_jitter * torch.randn(sd[k].shape, generator=g))
    return sd


def calibrate_running_stats(sd, x):
    """One training-mode oracle pass with momentum 1.0: running stats := batch
    stats of ``x`` (so eval-mode logits stay finite)."""
    from . import model_oracle as mo
    new = {}
    old = mo.BN_MOMENTUM
    mo.BN_MOMENTUM = 1.0
    try:
        with torch.no_grad():
            mo.forward(sd, x, training=True, new_stats=new)
    finally:
        mo.BN_MOMENTUM = old
    sd = dict(sd)
    sd.update(new)
    return sd


def synthetic_targets(B, H3, W3, seed=0, T=25, p_neg=0.98, p_pos=0.01):
    """class_map ~ {-1: p_neg, +1: p_pos, 0: rest} fp32 [B,T,H3,W3];
    regression_map ~ 0.2*randn [B,4T,H3,W3]."""
    r = np.random.RandomState(seed)
    u = r.rand(B, T, H3, W3)
    cm = np.zeros((B, T, H3, W3), np.float32)
    cm[u < p_neg] = -1
    cm[u > 1 - p_pos] = 1
    rm = (0.2 * r.randn(B, 4 * T, H3, W3)).astype(np.float32)
    return cm, rm


def synthetic_boxes(n, seed=0, extent=None, dup_frac=0.01):
    """NMS benchmark boxes: centres U(0,S)^2, sizes U(10,70)^2,
    scores U(0,1) plus a block of exact duplicates; float64."""
    r = np.random.RandomState(seed)
    if extent is None:
        extent = 40.0 * math.sqrt(n / 4.0)          # keeps the kept fraction roughly 10-30 %
    c = r.rand(n, 2) * extent
    wh = 10 + 60 * r.rand(n, 2)
    boxes = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
    scores = r.rand(n)
    nd = int(n * dup_frac)
    if nd:
        src = r.randint(0, n, nd)
        dst = r.randint(0, n, nd)
        boxes[dst] = boxes[src]
        scores[dst] = scores[src]
    return boxes.astype(np.float64), scores.astype(np.float64)


def load_templates():
    """The 25 templates the reference ships in tinyfaces/datasets/templates.json
    (consumed as constants); rounded copy kept in
    tests/golden/templates.json."""
    import json, os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "..", "tests", "golden", "templates.json")) as f:
        return np.array(json.load(f), dtype=np.float64)


RF = {"size": [859, 859], "stride": [8, 8], "offset": [-1, -1]}   # wider_face.py:55
