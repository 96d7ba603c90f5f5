"""numpy restatement of DetectionCriterion (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/tinyfaces/models/loss.py:59-93 and
/root/reference/tinyfaces/models/utils.py:103-163.  float32 arithmetic where
the reference uses float32 tensors; the balance sampler consumes
``np.random`` in exactly the reference's order (positives first, then
negatives, images in batch order).

Pinned by tests/test_oracle_golden.py against tests/golden/loss_*.npz.
"""
import numpy as np

OHEM_THRESH = 0.03      # loss.py:62
SAMPLE_SIZE = 256       # utils.py:103


def soft_margin(x, y):
    """nn.SoftMarginLoss(reduction='none'): log(1 + exp(-y*x)) in float32."""
    z = (-y * x).astype(np.float32)
    return np.log1p(np.exp(z, dtype=np.float32), dtype=np.float32)


def smooth_l1(d):
    """nn.SmoothL1Loss(reduction='none', beta=1) on a difference d (float32)."""
    a = np.abs(d)
    return np.where(a < 1, np.float32(0.5) * d * d, a - np.float32(0.5)).astype(np.float32)


def hard_negative_mining(cls, class_map):
    """loss.py:59-63 -- IN PLACE on class_map."""
    l = soft_margin(cls, class_map)
    class_map[l < np.float32(OHEM_THRESH)] = 0
    return class_map


def shuffle_index(n, n_out):
    """utils.py:142-163."""
    n, n_out = int(n), int(n_out)
    if n == 0 or n_out == 0:
        return np.empty(0)
    x = np.random.permutation(n)
    assert n_out <= n
    if n_out != n:
        x = x[:n_out]
    return x


def balance_sampling(label, pos_fraction, sample_size=SAMPLE_SIZE):
    """utils.py:103-139 -- IN PLACE on one image's label map [T,H,W]."""
    pos_max = sample_size * pos_fraction
    flat = label.reshape(-1)
    pos_idx = np.flatnonzero(flat == 1)
    if pos_idx.size > pos_max:
        didx = shuffle_index(pos_idx.size, pos_idx.size - pos_max)   # deleted prefix
        flat[pos_idx[didx]] = 0
    neg_max = pos_max * (1 - pos_fraction) / pos_fraction
    neg_idx = np.flatnonzero(flat == -1)
    if neg_idx.size > neg_max:
        ridx = shuffle_index(neg_idx.size, neg_max)                  # kept prefix
        d = np.delete(np.arange(neg_idx.size), ridx)
        flat[neg_idx[d]] = 0
    return label


def criterion(output, class_map, regression_map, n_templates=25, reg_weight=1.0,
              pos_fraction=0.5, alias_cpu=False):
    """loss.py:65-93.  Mutates class_map in place (OHEM), returns a dict with
    the total loss, the two masked sums, the final label map and
    d(total)/d(output) in closed form.

    ``alias_cpu``: on a CPU tensor ``class_map.cpu().numpy()`` (loss.py:49) is a
    *view*, so the reference's sampler also writes through to the caller's
    tensor; on a CUDA tensor it is a copy and only OHEM is visible.  The product
    follows the CUDA behaviour; the CPU golden needs alias_cpu=True."""
    T = n_templates
    cls = output[:, :T]
    reg = output[:, T:]
    hard_negative_mining(cls, class_map)                    # in place, loss.py:62
    lab = class_map.copy()                                  # .cpu().numpy() copy, loss.py:49
    for b in range(lab.shape[0]):
        balance_sampling(lab[b], pos_fraction)
    if alias_cpu:
        class_map[...] = lab
    cl = soft_margin(cls, lab)
    cmask = (lab != 0).astype(np.float32)
    masked_cls = cmask * cl
    d = (reg - regression_map).astype(np.float32)
    rl = smooth_l1(d)
    rmask = np.tile(lab > 0, (1, 4, 1, 1)).astype(np.float32)
    masked_reg = rmask * rl
    cls_sum = masked_cls.sum(dtype=np.float32)
    reg_sum = masked_reg.sum(dtype=np.float32)
    total = cls_sum + np.float32(reg_weight) * reg_sum
    # closed-form gradient: d/dx log(1+e^{-yx}) = -y * sigmoid(-y x)
    z = (-lab * cls).astype(np.float64)
    g_cls = (-lab * (1.0 / (1.0 + np.exp(-z))) * cmask).astype(np.float32)
    g_reg = (np.clip(d, -1, 1) * rmask * np.float32(reg_weight)).astype(np.float32)
    grad = np.concatenate([g_cls, g_reg], axis=1)
    return dict(total=total, cls_sum=cls_sum, reg_sum=reg_sum, labels=lab, grad=grad,
                masked_cls=masked_cls, masked_reg=masked_reg)
