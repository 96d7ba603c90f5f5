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


def state_dict(seed=0, num_templates=25, bn3_gamma=1.0, beta_jitter=0.0):
    """Seeded synthetic weights with the reference's 571 state_dict keys (model.py:12-40 + torchvision resnet101 minus
    layer4): convs ~ kaiming-normal(fan_out), BN gamma 1 (bn3 gamma configurable), heads ~ U(+-1/sqrt(fan_in)), the frozen
    bilinear upsample kernel.  Draw-for-draw the recipe the golden fixtures were generated with (tests/test_cabi_cpu.py
    checks it against the test-side generator), so bench.py can score its output against tests/golden/cfg2_b8_fwd.npz."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(cout, cin, k):
        return torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cout * k * k))

    def bn(prefix, c, gamma=1.0):
        sd[prefix + ".weight"] = torch.full((c,), float(gamma))
        sd[prefix + ".bias"] = torch.zeros(c)
        sd[prefix + ".running_mean"] = torch.zeros(c)
        sd[prefix + ".running_var"] = torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    sd["model.conv1.weight"] = conv(64, 3, 7)
    bn("model.bn1", 64)
    inplanes = 64
    for name, blocks, planes in (("layer1", 3, 64), ("layer2", 4, 128), ("layer3", 23, 256)):
        for i in range(blocks):
            p = "model.%s.%d" % (name, i)
            sd[p + ".conv1.weight"] = conv(planes, inplanes, 1)
            bn(p + ".bn1", planes)
            sd[p + ".conv2.weight"] = conv(planes, planes, 3)
            bn(p + ".bn2", planes)
            sd[p + ".conv3.weight"] = conv(planes * 4, planes, 1)
            bn(p + ".bn3", planes * 4, bn3_gamma)
            if i == 0:
                sd[p + ".downsample.0.weight"] = conv(planes * 4, inplanes, 1)
                bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    sd["model.fc.weight"] = (torch.rand(1000, 2048, generator=g) * 2 - 1) / math.sqrt(2048)
    sd["model.fc.bias"] = (torch.rand(1000, generator=g) * 2 - 1) / math.sqrt(2048)
    out = 5 * num_templates
    for nm, cin in (("score_res3", 512), ("score_res4", 1024)):
        b = 1.0 / math.sqrt(cin)
        sd[nm + ".weight"] = (torch.rand(out, cin, 1, 1, generator=g) * 2 - 1) * b
        sd[nm + ".bias"] = (torch.rand(out, generator=g) * 2 - 1) * b
    taps = torch.tensor([0.25, 0.75, 0.75, 0.25], dtype=torch.float32)          # model.py:45-65
    w = torch.zeros(out, out, 4, 4)
    ch = torch.arange(out)
    w[ch, ch] = torch.outer(taps, taps)
    sd["score4_upsample.weight"] = w
    if beta_jitter:
        for k in list(sd):
            if k.endswith(".bias") and ("bn" in k or "downsample.1" in k):
                sd[k] = torch.randn(sd[k].shape, generator=g) * beta_jitter
            if k.endswith(".weight") and sd[k].dim() == 1:
                sd[k] = sd[k] * (1 + beta_jitter * torch.randn(sd[k].shape, generator=g))
    return sd
