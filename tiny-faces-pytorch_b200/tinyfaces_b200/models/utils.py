"""get_bboxes / balance_sampling -- drop-ins for /root/reference/tinyfaces/models/utils.py.

``get_bboxes`` keeps the reference's numpy-in / numpy-out signature (utils.py:4-11) but thresholds, compacts
(in the reference's (b, y, x, c) order) and decodes on the GPU through ``tf_decode``.
``balance_sampling`` / ``shuffle_index`` are the host-side sampler that consumes ``np.random`` exactly like the
reference (utils.py:103-163) -- needed for bit-identical label maps; the fast path is the device sampler
(``tf_detloss_sample_device``).
"""
import numpy as np
import torch

from .. import ops


def invalid_template_ids(templates, scale):
    """utils.py:17-41: templates that must not fire at this pyramid scale."""
    every_scale = np.arange(4, 12)                 # type-A templates
    single_scale = np.arange(18, 25)               # type-B templates
    never = np.setdiff1d(np.arange(25), np.concatenate((every_scale, single_scale)))
    tscale = np.asarray(templates)[:, 4][single_scale]
    wrong = (tscale >= 1.0) if scale < 1 else (tscale != 1.0)
    return np.concatenate((never, single_scale[wrong]))


def _bitmask(ids):
    m = 0
    for i in ids:
        m |= 1 << int(i)
    return m


def get_bboxes(score_cls, score_reg, prob_cls, templates, prob_thresh, rf, scale=1, refine=True, bug_compat=True,
               device=None):
    """utils.py:4-76.  NHWC float32 numpy maps -> (bboxes float64 [N,4], scores float32 [N,1]).

    ``bug_compat=True`` (default) reproduces the shipped behaviour of utils.py:44: the invalid template ids index
    the *x* axis of the NHWC probability map (and the input ``prob_cls`` is modified in place, as in the reference).
    ``bug_compat=False`` masks templates instead.
    """
    if not refine:
        raise NotImplementedError("refine=False is a dead path in the reference (utils.py:66-70)")
    templates = np.asarray(templates, dtype=np.float64)
    T = templates.shape[0]
    inv = invalid_template_ids(templates, scale)
    if bug_compat:
        prob_cls[:, :, inv] = 0.0              # same in-place side effect (and the same IndexError when W < 25)
    else:
        prob_cls[:, :, :, inv] = 0.0
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    sc = torch.from_numpy(np.ascontiguousarray(score_cls, dtype=np.float32)).to(dev)
    sr = torch.from_numpy(np.ascontiguousarray(score_reg, dtype=np.float32)).to(dev)
    pr = torch.from_numpy(np.ascontiguousarray(prob_cls, dtype=np.float32)).to(dev)
    B, H, W, _ = sc.shape
    boxes, scores, _, count = ops.decode_device(
        sc, sr, pr, (H * W * T, W * T, T, 1), (H * W * 4 * T, W * 4 * T, 4 * T, 1), B, H, W, T, templates, prob_thresh,
        0, 0, rf, scale)                        # masks already applied to prob_cls above
    n = int(count.item())
    return boxes[:n].cpu().numpy(), scores[:n].to(torch.float32).cpu().numpy().reshape(n, 1)


def shuffle_index(n, n_out):
    """utils.py:142-163: the first n_out entries of one np.random.permutation(n) draw."""
    n, n_out = int(n), int(n_out)
    if n == 0 or n_out == 0:
        return np.empty(0)
    perm = np.random.permutation(n)
    assert n_out <= n
    return perm if n_out == n else perm[:n_out]


def balance_sampling(label_cls, pos_fraction, sample_size=256):
    """utils.py:103-139, in place on one image's [T,H,W] label map.

    RNG protocol (must match the reference draw for draw): positives first -- only when there are more than
    ``sample_size*pos_fraction`` -- where the permuted prefix is the set that is *removed*; then negatives, where
    the permuted prefix is the set that is *kept*."""
    flat = label_cls.reshape(-1)
    max_pos = sample_size * pos_fraction
    pos = np.flatnonzero(flat == 1)
    if pos.size > max_pos:
        drop = shuffle_index(pos.size, pos.size - max_pos)
        flat[pos[drop]] = 0
    max_neg = max_pos * (1 - pos_fraction) / pos_fraction
    neg = np.flatnonzero(flat == -1)
    if neg.size > max_neg:
        keep = shuffle_index(neg.size, max_neg)
        mask = np.ones(neg.size, dtype=bool)
        mask[keep] = False
        flat[neg[mask]] = 0
    return label_cls
