"""numpy restatement of get_bboxes / regression_refinement (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/tinyfaces/models/utils.py:4-100, including the
shipped quirk at :44 (the invalid-template ids index the *W* axis of the
NHWC probability map) and its in-place mutation of
``prob_cls``.  Written with explicit flat-index arithmetic rather than
fancy indexing so it shares no code shape with the reference.

Pinned by tests/test_oracle_golden.py against tests/golden/decode_*.npz.
"""
import numpy as np


def invalid_ids(templates, scale):
    """utils.py:18-41."""
    all_scale = np.arange(4, 12)
    one_scale = np.arange(18, 25)
    ignored = np.setdiff1d(np.arange(25), np.concatenate((all_scale, one_scale)))
    ts = templates[:, 4][one_scale]
    bad = ts >= 1.0 if scale < 1 else ts != 1.0
    return np.concatenate((ignored, one_scale[bad]))


def get_bboxes(score_cls, score_reg, prob_cls, templates, prob_thresh, rf, scale=1):
    """utils.py:4-76 with refine=True.  NHWC float32 maps -> (boxes f64 [N,4], scores f32 [N,1])."""
    T = templates.shape[0]
    B, H, W, _ = prob_cls.shape
    inv = invalid_ids(templates, scale)
    prob_cls[:, :, inv] = 0.0                                  # utils.py:44 (axis 2 == x)
    flat = np.flatnonzero(prob_cls.reshape(-1) > prob_thresh)  # C order == (b,y,x,c)
    fc = flat % T
    fx = (flat // T) % W
    fy = (flat // (T * W)) % H
    fb = flat // (T * W * H)
    scores = score_cls.reshape(-1)[flat].reshape(-1, 1)
    stride, offset = rf["stride"], rf["offset"]
    cy = fy * stride[0] + offset[0]
    cx = fx * stride[1] + offset[1]
    cw = templates[fc, 2] - templates[fc, 0] + 1
    ch = templates[fc, 3] - templates[fc, 1] + 1
    pix = ((fb * H + fy) * W + fx) * (4 * T)
    reg = score_reg.reshape(-1)
    tx, ty = reg[pix + fc], reg[pix + T + fc]
    tw, th = reg[pix + 2 * T + fc], reg[pix + 3 * T + fc]
    rcx = cx + cw * tx                                         # utils.py:81-85
    rcy = cy + ch * ty
    rcw = cw * np.exp(tw)                                      # float32 exp, f64 product (:87-88)
    rch = ch * np.exp(th)
    boxes = np.stack([rcx - rcw / 2, rcy - rch / 2, rcx + rcw / 2, rcy + rch / 2], axis=1)
    boxes = boxes * (1 / scale)                                # :73-74
    return boxes.reshape(-1, 4), scores
