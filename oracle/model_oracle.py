"""CPU fp32 restatement of DetectionModel.forward (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/tinyfaces/models/model.py:89-128 (forward) and the
torchvision ResNet-101 trunk it instantiates (torchvision/models/resnet.py:
108-163 Bottleneck v1.5 -- stride on the 3x3 --, :197-204 stem).  It is
*functional*: it takes the reference's ``state_dict`` (same key names) and
calls plain ``torch.nn.functional`` CPU ops, so it shares no module code
with either the reference or the product.

Pinned by tests/test_oracle_golden.py against tests/golden/model_*.npz, which
oracle/make_golden.py produced by running the reference itself.
"""
import numpy as np
import torch
import torch.nn.functional as F

LAYERS = (("layer1", 3, 64, 1), ("layer2", 4, 128, 2), ("layer3", 23, 256, 2))
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def bilinear_upsample_weight(channels, k=4):
    """model.py:45-65 -- diagonal bilinear ConvTranspose2d weight [C,C,k,k]."""
    factor = np.floor((k + 1) / 2)
    center = factor if k % 2 == 1 else factor + 0.5
    c = np.arange(1, k + 1)
    v = (np.ones((1, k)) - (np.abs(c - center) / factor))
    f2 = v.T @ v
    w = np.zeros((channels, channels, k, k))
    for i in range(channels):
        w[i, i] = f2
    return torch.tensor(w, dtype=torch.float32)


def _bn(x, sd, prefix, training, new_stats):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm2, rv2, w, b, True, BN_MOMENTUM, BN_EPS)
        if new_stats is not None:
            new_stats[prefix + ".running_mean"] = rm2
            new_stats[prefix + ".running_var"] = rv2
        return y
    return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)


def _bottleneck(x, sd, p, stride, training, new_stats, taps):
    out = F.conv2d(x, sd[p + ".conv1.weight"])
    out = F.relu(_bn(out, sd, p + ".bn1", training, new_stats))
    out = F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)
    out = F.relu(_bn(out, sd, p + ".bn2", training, new_stats))
    out = F.conv2d(out, sd[p + ".conv3.weight"])
    out = _bn(out, sd, p + ".bn3", training, new_stats)
    if (p + ".downsample.0.weight") in sd:
        idt = F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride)
        idt = _bn(idt, sd, p + ".downsample.1", training, new_stats)
    else:
        idt = x
    out = F.relu(out + idt)
    if taps is not None:
        taps[p] = out
    return out


def forward(sd, x, training=False, new_stats=None, taps=None):
    """model.py:89-128.  ``sd``: reference state_dict (fp32 CPU tensors).

    ``new_stats`` (dict) receives the updated BN running statistics in
    training mode; ``taps`` (dict) receives intermediate activations.
    """
    x = F.conv2d(x, sd["model.conv1.weight"], stride=2, padding=3)   # :90
    x = F.relu(_bn(x, sd, "model.bn1", training, new_stats))          # :91-92
    x = F.max_pool2d(x, 3, 2, 1)                                      # :93
    if taps is not None:
        taps["stem"] = x
    feats = {}
    for name, blocks, _planes, stride in LAYERS:                      # :95-101
        for i in range(blocks):
            x = _bottleneck(x, sd, "model.%s.%d" % (name, i),
                            stride if i == 0 else 1, training, new_stats, taps)
        feats[name] = x
    res3, res4 = feats["layer2"], feats["layer3"]
    s3 = F.conv2d(res3, sd["score_res3.weight"], sd["score_res3.bias"])   # :104
    s4 = F.conv2d(res4, sd["score_res4.weight"], sd["score_res4.bias"])   # :106
    up = F.conv_transpose2d(s4, sd["score4_upsample.weight"], stride=2, padding=1)  # :107
    # both crop branches (:110-124) keep the first H3 rows / W3 columns
    up = up[:, :, :s3.shape[2], :s3.shape[3]]
    return s3 + up                                                    # :126
