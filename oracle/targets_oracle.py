"""CPU restatement of the reference's training-target generation (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/tinyfaces/datasets/processor.py:223-277 (get_heatmaps), :157-221 (get_regression) and
/root/reference/tinyfaces/datasets/dense_overlap.py:4-75 (compute_dense_overlap), vectorised with numpy broadcasting
instead of the reference's 4-deep Python loop; every float64 operation keeps the reference's order.  It consumes
np.random exactly like the reference (one np.random.rand(vsy, vsx, nt, ng) call, processor.py:203).

Pinned by tests/test_oracle_golden.py against tests/golden/targets_case*.npz, which oracle/make_golden.py produced by
running the reference's DataProcessor itself (its clustering-only imports stubbed out).
"""
import numpy as np


def dense_overlap(ofx, ofy, stx, sty, vsx, vsy, tpl, boxes):
    """dense_overlap.py:4-75 with zmx = zmy = 1 -> [vsy, vsx, nt, ng], rounded to 14 decimals."""
    cx = (ofx + np.arange(vsx) * (stx / 1))[None, :, None, None]
    cy = (ofy + np.arange(vsy) * (sty / 1))[:, None, None, None]
    dx1, dy1, dx2, dy2 = (tpl[:, k][None, None, :, None] for k in range(4))
    gx1, gy1, gx2, gy2 = (boxes[:, k][None, None, None, :] for k in range(4))
    barea = (gx2 - gx1 + 1) * (gy2 - gy1 + 1)
    farea = (dx2 - dx1 + 1) * (dy2 - dy1 + 1)
    x1, y1, x2, y2 = dx1 + cx, dy1 + cy, dx2 + cx, dy2 + cy
    iw = np.minimum(x2, gx2) - np.maximum(x1, gx1) + 1
    ih = np.minimum(y2, gy2) - np.maximum(y1, gy1) + 1
    ia = iw * ih
    with np.errstate(invalid="ignore", divide="ignore"):
        ov = np.where((ih > 0) & (iw > 0), ia / (farea + barea - ia), 0.0)
    return np.around(ov, decimals=14)


def get_heatmaps(bboxes, pad_mask, templates, rf, heatmap_size, pos_thresh, neg_thresh):
    ofy, ofx = rf["offset"]
    sty, stx = rf["stride"]
    vsy, vsx = heatmap_size
    tpl = np.asarray(templates, dtype=np.float64)[:, :4]
    nt = tpl.shape[0]
    class_maps = -np.ones((vsy, vsx, nt))
    regress_maps = np.zeros((vsy, vsx, nt * 4))
    bboxes = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
    invalid = np.logical_or(bboxes[:, 2] <= bboxes[:, 0], bboxes[:, 3] <= bboxes[:, 1])      # processor.py:236-240
    bboxes = np.delete(bboxes, np.where(invalid), axis=0)
    ng = bboxes.shape[0]
    iou = np.zeros((vsy, vsx, nt, ng))
    if ng > 0:
        iou = dense_overlap(ofx, ofy, stx, sty, vsx, vsy, tpl, bboxes)
        # ---- get_regression (processor.py:157-221)
        cxx = (ofx + np.arange(vsx) * stx)[None, :, None, None]
        cyy = (ofy + np.arange(vsy) * sty)[:, None, None, None]
        dww = (tpl[:, 2] - tpl[:, 0] + 1)[None, None, :, None]
        dhh = (tpl[:, 3] - tpl[:, 1] + 1)[None, None, :, None]
        f = [bboxes[:, k][None, None, None, :] for k in range(4)]
        tx = np.divide((f[0] + f[2]) / 2 - cxx, dww)
        ty = np.divide((f[1] + f[3]) / 2 - cyy, dhh)
        tw = np.broadcast_to(np.log(np.divide(f[2] - f[0] + 1, dww)), iou.shape)
        th = np.broadcast_to(np.log(np.divide(f[3] - f[1] + 1, dhh)), iou.shape)
        iou = iou + (1e-6 * np.random.rand(*iou.shape))                                       # processor.py:203
        best = iou.argmax(axis=3)[..., None]
        pick = lambda a: np.take_along_axis(np.broadcast_to(a, iou.shape), best, axis=3)[..., 0]   # noqa: E731
        regress_maps = np.concatenate((pick(tx), pick(ty), pick(tw), pick(th)), axis=2)
        best_iou = iou.max(axis=3)
        per_obj = iou.reshape(-1, ng)                                                          # processor.py:250-256
        fbest = np.argmax(per_obj, axis=0)
        keep = np.amax(per_obj, axis=0) > neg_thresh
        class_maps[np.unravel_index(fbest[keep], iou.shape[:-1])] = 1
        class_maps = np.maximum(class_maps, (best_iou >= pos_thresh) * 2 - 1)
        gray = -np.ones(class_maps.shape)
        gray[np.bitwise_and(neg_thresh <= best_iou, best_iou < pos_thresh)] = 0
        class_maps = np.maximum(class_maps, gray)
    non_neg_border = np.bitwise_and(pad_mask, class_maps != -1)                                # processor.py:271-273
    class_maps[non_neg_border] = 0
    regress_maps[:, :, :nt][non_neg_border] = 0
    return class_maps, regress_maps, iou
