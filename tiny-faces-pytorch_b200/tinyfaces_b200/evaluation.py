"""get_detections / nms -- drop-ins for /root/reference/tinyfaces/evaluation.py:20-87 and the
``torchvision.ops.nms`` call it makes (evaluation.py:84).

Per pyramid level only the resized image goes up and nothing comes down: forward, sigmoid + threshold +
order-preserving compaction + anchor decode (``tf_decode``) and the final global NMS (``tf_nms``) stay on the GPU;
the K kept boxes are the only device->host copy.  (The reference copies 150 floats per heat-map pixel per level
to the host, evaluation.py:64-68, and runs NMS single-threaded on the CPU.)
"""
import numpy as np
import torch
from torchvision import transforms

from . import ops
from .models.utils import _bitmask, invalid_template_ids
from .pyramid import pyramid_level


def _normalize_params(img_transforms):
    """(mean, std) if ``img_transforms`` is the reference's Compose([ToTensor(), Normalize(mean, std)])
    (evaluate_model.py / detect_image.py), else None -> the PIL host path is used."""
    ts = getattr(img_transforms, "transforms", None)
    if ts is not None and len(ts) == 2 and isinstance(ts[0], transforms.ToTensor) and isinstance(ts[1], transforms.Normalize):
        return [float(v) for v in ts[1].mean], [float(v) for v in ts[1].std]
    return None


class _Pyramid:
    """Produces the per-scale network input exactly as evaluation.py:40-56 does.  With the standard transforms the
    whole level (uint8 quantisation, PIL-exact bilinear resize, ToTensor, Normalize) is built on the GPU
    (tf_pyramid_level); any other transform falls back to the reference's own PIL call sequence on the host."""

    def __init__(self, img, img_transforms, device, gpu_pyramid=True):
        self.device = device
        self.tf = img_transforms
        self.params = _normalize_params(img_transforms) if gpu_pyramid else None
        if self.params is not None:
            self.img = img.to(device=device, dtype=torch.float32).contiguous()
            self.min_side = min(img.shape[1], img.shape[2])
        else:
            self.image = transforms.functional.to_pil_image(img)            # evaluation.py:40
            self.min_side = np.min(self.image.size)

    def level(self, scale):
        size = int(self.min_side * scale)                                   # evaluation.py:46-47
        if self.params is not None:
            return pyramid_level(self.img, size, *self.params)
        scaled = transforms.functional.resize(self.image, size)
        return self.tf(scaled).unsqueeze(0).float().to(self.device, non_blocking=True)


def _to_device(model, device):
    """model.to(device) (evaluation.py:33), skipped when the parameters already live there: nn.Module.to walks every
    sub-module and drops the executor's cached pointer tables even when nothing moves."""
    p = next(model.parameters(), None)
    if p is not None and p.device == device:
        return model
    return model.to(device)


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms(boxes[N,4], scores[N], iou_threshold) -> int64[K] (descending score), bit-identical
    keep indices.  CPU inputs are moved to the current CUDA device and the result is returned on the input's
    device, like the op it replaces."""
    if boxes.dtype != scores.dtype:
        raise RuntimeError("nms: boxes and scores should have the same dtype")
    src = boxes.device
    dev = src if boxes.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return ops.nms_keep(boxes.to(dev), scores.to(dev), iou_threshold).to(src)


def decode_level(output, templates, prob_thresh, rf, scale, bug_compat=True, sync=True, rows=None, row_offset=0):
    """Device-side replacement of evaluation.py:61-71 for one pyramid level.
    output: [B,5T,H,W] CUDA tensor.  Returns (boxes f64 [N,4], scores f64 [N]) on the device; with sync=False the
    capacity-sized buffers and the device-side count are returned instead (boxes, scores, count) so that the caller can
    read all counts with ONE host synchronisation after every level has been enqueued.
    rows=(a, b) decodes only heat-map rows [a, b) of `output`, and row_offset is the GLOBAL heat-map row of output row 0
    (spatially tiled levels: the tile's rows are decoded as the rows of the whole level they are)."""
    B, C, H, W = output.shape
    T = templates.shape[0]
    if bug_compat and W < 25:
        raise IndexError("index 24 is out of bounds for axis 2 with size %d" % W)   # the reference's failure mode
    inv = _bitmask(invalid_template_ids(templates, scale))
    hw = H * W
    strides = (C * hw, W, 1, hw)                       # (b, y, x, c) element strides of an NCHW tensor
    out = output.contiguous()
    a, b = (0, H) if rows is None else rows
    if row_offset + a:
        # the kernel numbers the rows of the VIEW from 0: view row j is global heat-map row row_offset + a + j.
        # cy = y * stride + offset is integer arithmetic (utils.py:53): shifting the offset by whole rows is exact
        rf = dict(rf, offset=[int(rf["offset"][0]) + int(rf["stride"][0]) * int(row_offset + a), int(rf["offset"][1])])
    view = out[:, :, a:b, :]
    boxes, scores, _, count = ops.decode_device(view, view[:, T:], None, strides, strides, B, b - a, W, T, templates,
                                                prob_thresh, inv if bug_compat else 0, 0 if bug_compat else inv, rf, scale)
    if not sync:
        return boxes, scores, count
    n = int(count.item())
    return boxes[:n], scores[:n]


# ------------------------------------------------------------------------------------------------ spatial tiling
# A pyramid level can be cut into horizontal bands that are evaluated independently (on different GPUs): eval-mode BN is
# an affine map, so a heat-map row depends only on the input rows inside its receptive field.  Every band carries a halo
# of HALO_PX input rows on each cut side -- at least half the theoretical receptive field (859 px,
# tinyfaces/datasets/wider_face.py:55) -- and starts on a multiple of 16 input rows (the trunk's total stride at res4, so
# the stride-2 phases of the tile coincide with those of the whole image).  Inside the band the tile's heat-map rows are
# then BIT-identical to the rows of the whole-level forward (the inference GEMMs have a fixed K order per output pixel,
# zero padding == TMA out-of-bounds fill).  Bands are full-width, so concatenating their candidates in band order is the
# level's (y, x, c) order -- the global NMS sees exactly the single-pass candidate list.
HALO_PX = 448


def out_rows(h):
    """heat-map rows of an input with h rows (three stride-2 stages with ceil)"""
    for _ in range(3):
        h = (h - 1) // 2 + 1
    return h


def plan_bands(H, nbands):
    """Cut the out_rows(H) heat-map rows of a level into `nbands` bands.  Returns [(r0, r1, y0, y1)]: heat-map rows
    [r0, r1) come from the tile of input rows [y0, y1)."""
    H3 = out_rows(H)
    nbands = max(1, min(int(nbands), H3 // 2))
    # equal TILE heights, not equal band heights: the two edge bands carry one halo, the inner ones two, so the edge bands
    # get HALO_PX more interior rows (tile height t solves (nbands - 2) (t - 2 halo) + 2 (t - halo) = H)
    t = (H + 2 * HALO_PX * (nbands - 1)) / float(nbands)
    edge, inner = t - HALO_PX, t - 2 * HALO_PX
    cuts = [0]
    for k in range(1, nbands):
        px = edge + (k - 1) * inner if inner > 64 else H * k / float(nbands)
        r = int(px / 8.0) // 2 * 2                      # even heat-map rows = multiples of 16 input rows
        if cuts[-1] < r < H3:
            cuts.append(r)
    cuts.append(H3)
    bands = []
    for r0, r1 in zip(cuts, cuts[1:]):
        y0 = max(0, (8 * r0 - HALO_PX) // 16 * 16)
        y1 = H if r1 == H3 else min(H, 8 * r1 + HALO_PX)
        bands.append((r0, r1, y0, y1))
    return bands


_PLAN_CACHE = {}
# forward time of one tile on a B200 ~ FIXED + PER_MPIX * megapixels (measured: 312 px level 1.9 ms, 5000 px level 23.5 ms):
# the fixed part (~100 launch-bound kernels) is what makes a pure pixel count over-value many small jobs on one rank
JOB_FIXED_MS, JOB_MS_PER_MPIX = 1.7, 0.9


def plan_jobs(level_shapes, world, spatial=True, force_bands=None):
    """Work list for `world` ranks: [(level, band, r0, r1, y0, y1, owner)].  Levels are cut into bands only when that
    lowers the makespan of a longest-processing-time packing (cost = estimated forward time of the tile, halo
    included); the band counts of the two largest levels are searched exhaustively.  Plans are cached."""
    key = (tuple(map(tuple, level_shapes)), world, bool(spatial), tuple(sorted(force_bands.items())) if force_bands else None)
    if key in _PLAN_CACHE:
        return _PLAN_CACHE[key]

    def pack(band_counts):
        jobs = []
        for lv, (H, W) in enumerate(level_shapes):
            for bi, (r0, r1, y0, y1) in enumerate(plan_bands(H, band_counts.get(lv, 1))):
                jobs.append([lv, bi, r0, r1, y0, y1, JOB_FIXED_MS + JOB_MS_PER_MPIX * (y1 - y0) * W / 1e6])
        load = [0.0] * world
        for j in sorted(jobs, key=lambda j: -j[6]):
            r = min(range(world), key=lambda k: load[k])
            j.append(r)
            load[r] += j[6]
        return max(load), jobs
    order = sorted(range(len(level_shapes)), key=lambda i: -level_shapes[i][0] * level_shapes[i][1])
    best = pack(dict(force_bands) if force_bands else {})
    if force_bands is None and spatial and world > 1 and order:
        big, second = order[0], (order[1] if len(order) > 1 else None)
        for nb in range(1, 2 * world + 1):
            for nb2 in (range(1, world + 1) if second is not None else [1]):
                cand = pack({big: nb, **({second: nb2} if second is not None else {})})
                if cand[0] < best[0] * 0.999:
                    best = cand
    jobs = sorted(best[1], key=lambda j: (j[0], j[1]))
    plan = [(lv, bi, r0, r1, y0, y1, owner) for lv, bi, r0, r1, y0, y1, _c, owner in jobs]
    _PLAN_CACHE[key] = plan
    return plan


def run_band(model, x_level, job, templates, prob_thresh, rf, scale):
    """Forward + decode of one band of one level (x_level: the whole level image [1,3,H,W] on the device).
    Returns the unsynchronised (boxes, scores, count) of decode_level."""
    _lv, _bi, r0, r1, y0, y1, _owner = job
    H = x_level.shape[2]
    xt = x_level if (y0 == 0 and y1 == H) else x_level[:, :, y0:y1, :].contiguous()
    with torch.no_grad():
        out = model(xt)
    a = r0 - y0 // 8
    return decode_level(out, templates, prob_thresh, rf, scale, sync=False, rows=(a, a + (r1 - r0)), row_offset=y0 // 8)


def get_detections_tiled(model, img, templates, rf, img_transforms, prob_thresh=0.65, nms_thresh=0.3, scales=(-2, -1, 0, 1),
                         device=None, bands=2, return_candidates=False):
    """get_detections with every level cut into `bands` bands (an int or {level index: count}), all on ONE GPU: the
    single-device proof that spatial tiling reproduces the untiled result bit for bit."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    model = _to_device(model, device)
    model.eval()
    templates = np.asarray(templates, dtype=np.float64)
    pyr = _Pyramid(img, img_transforms, device)
    parts = []
    for lv, scale in enumerate([2 ** x for x in scales]):
        x = pyr.level(scale)
        nb = bands.get(lv, 1) if isinstance(bands, dict) else bands
        for bi, (r0, r1, y0, y1) in enumerate(plan_bands(x.shape[2], nb)):
            parts.append(run_band(model, x, (lv, bi, r0, r1, y0, y1, 0), templates, prob_thresh, rf, scale))
    counts = torch.cat([c for _b, _s, c in parts]).cpu().tolist()
    boxes = torch.cat([b[:n] for (b, _s, _c), n in zip(parts, counts)])
    scores = torch.cat([s[:n] for (_b, s, _c), n in zip(parts, counts)])
    if return_candidates:
        return boxes, scores
    return boxes[ops.nms_keep(boxes, scores, nms_thresh)].cpu().numpy()


def get_detections(model, img, templates, rf, img_transforms, prob_thresh=0.65, nms_thresh=0.3, scales=(-2, -1, 0, 1),
                   device=None, return_scores=False, gpu_pyramid=True):
    """evaluation.py:20-87.  img: CHW float tensor in [0,1]; scales are exponents of 2; returns ndarray [K,4] float64."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    model = _to_device(model, device)
    model.eval()
    templates = np.asarray(templates, dtype=np.float64)
    pyr = _Pyramid(img, img_transforms, device, gpu_pyramid)
    levels = []
    for scale in [2 ** x for x in scales]:                          # evaluation.py:37,44
        x = pyr.level(scale)
        with torch.no_grad():
            output = model(x)
        levels.append(decode_level(output, templates, prob_thresh, rf, scale, sync=False))
    # one host synchronisation for all levels: the host has enqueued every level before it waits for any candidate count
    counts = torch.cat([c for _b, _s, c in levels]).cpu().tolist() if levels else []
    all_boxes = [b[:n] for (b, _s, _c), n in zip(levels, counts)]
    all_scores = [s[:n] for (_b, s, _c), n in zip(levels, counts)]
    boxes = torch.cat(all_boxes) if all_boxes else torch.zeros((0, 4), dtype=torch.float64, device=device)
    scores = torch.cat(all_scores) if all_scores else torch.zeros(0, dtype=torch.float64, device=device)
    keep = ops.nms_keep(boxes, scores, nms_thresh)                  # evaluation.py:84
    dets = boxes[keep].cpu().numpy()                                # evaluation.py:85-87
    if return_scores:
        return dets, scores[keep].cpu().numpy()
    return dets


def write_results(dets, img_path, split, results_dir=None, scores=None):
    """evaluation.py:89-114 (WIDER-FACE result file: "left top width height score" per detection), made to work: the
    shipped function reads the score from column 4 of `dets`, but its get_detections returns [K,4] boxes only (SURVEY
    App. A), so evaluate_model.py:58-67 raises on the first detection.  Here `dets` may be [K,5] (boxes + score, as the
    file format intends) or [K,4] with `scores` [K] passed separately -- e.g. get_detections(..., return_scores=True)."""
    from pathlib import Path
    dets = np.asarray(dets)
    if dets.ndim != 2 or dets.shape[1] not in (4, 5):
        raise ValueError("write_results: dets must be [K,4] or [K,5]")
    if dets.shape[1] == 4:
        if scores is None:
            raise ValueError("write_results: [K,4] detections need `scores` (get_detections(..., return_scores=True))")
        dets = np.concatenate([dets, np.asarray(scores, dtype=np.float64).reshape(-1, 1)], axis=1)
    results_dir = Path(results_dir or f"{split}_results")
    filename = results_dir / img_path.replace('jpg', 'txt')
    filename.parent.mkdir(parents=True, exist_ok=True)
    with open(filename, 'w') as f:
        f.write(img_path.split('/')[-1] + "\n")
        f.write(str(dets.shape[0]) + "\n")
        for x in dets:
            left, top = np.round(x[0]), np.round(x[1])
            width, height = np.round(x[2] - x[0] + 1), np.round(x[3] - x[1] + 1)
            f.write(f"{int(left)} {int(top)} {int(width)} {int(height)} {x[4]}\n")
    return filename


# ------------------------------------------------------------------------------------------------ multi-GPU
def gather_level_candidates(per_level, num_levels, group=None, dst=0, device=None):
    """Scale-sharded inference exchange step.  ``per_level``: {level_index: (boxes [n,4] f64, scores [n] f64)} for
    the pyramid levels this rank evaluated.  Every rank contributes its levels; rank ``dst`` receives all
    candidates concatenated in level order (the reference's ``scales`` order, evaluation.py:78) -- which is what
    makes the global NMS keep indices identical to the single-GPU run.  Returns (boxes, scores) on ``dst`` and
    (None, None) elsewhere.  Variable sizes are exchanged first, payloads are padded to the maximum."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    # `device`: where this rank's collective buffers live.  A rank that owns no level (more ranks than levels) has no
    # tensor to infer it from -- under NCCL a CPU tensor there would hang the collective.
    any_t = next(iter(per_level.values()))[0] if per_level else None
    if device is not None:
        dev = torch.device(device)
    elif any_t is not None:
        dev = any_t.device
    elif dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
    counts = torch.zeros(num_levels, dtype=torch.int64, device=dev)
    for lv, (b, _s) in per_level.items():
        counts[lv] = b.shape[0]
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    table = torch.stack(all_counts).cpu().tolist()               # [rank][level] candidate counts: ONE host read
    mine = sorted(per_level)
    payload = torch.cat([torch.cat([per_level[lv][0], per_level[lv][1][:, None]], dim=1) for lv in mine]) \
        if mine else torch.zeros((0, 5), dtype=torch.float64, device=dev)
    sizes = [sum(row) for row in table]
    cap = max(max(sizes), 1)
    padded = torch.zeros((cap, 5), dtype=torch.float64, device=dev)
    padded[: payload.shape[0]] = payload
    gathered = [torch.zeros_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, gathered, dst=dst, group=group)
    if rank != dst:
        return None, None
    chunks = [None] * num_levels
    for r in range(world):
        off = 0
        for lv in range(num_levels):
            n = table[r][lv]
            if n:
                chunks[lv] = gathered[r][off:off + n]
                off += n
    cat = torch.cat([c for c in chunks if c is not None]) if any(c is not None for c in chunks) \
        else torch.zeros((0, 5), dtype=torch.float64, device=dev)
    assert cat.shape[0] == sum(sizes)
    return cat[:, :4].contiguous(), cat[:, 4].contiguous()


def get_detections_sharded(model, img, templates, rf, img_transforms, prob_thresh=0.65, nms_thresh=0.3,
                           scales=(-2, -1, 0, 1), device=None, group=None, spatial=True, return_plan=False, force_bands=None):
    """get_detections sharded over the ranks of `group`: the pyramid levels -- and, with spatial=True, horizontal bands of
    the large levels (plan_jobs) -- are packed onto the ranks by cost, every rank evaluates its jobs, the candidates are
    gathered to rank 0 in (level, band) order, i.e. the single-GPU candidate order, and the global NMS runs there.
    Rank 0 returns ndarray [K,4] (identical to get_detections); other ranks return None."""
    import torch.distributed as dist
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    model = _to_device(model, device)
    model.eval()
    templates = np.asarray(templates, dtype=np.float64)
    pyr = _Pyramid(img, img_transforms, device)
    levels = [2 ** x for x in scales]
    H0, W0 = int(img.shape[1]), int(img.shape[2])
    from .pyramid import resized_size
    shapes = []
    for sc in levels:
        wo, ho = resized_size(W0, H0, int(min(H0, W0) * sc))
        shapes.append((ho, wo))
    jobs = plan_jobs(shapes, world, spatial, force_bands)        # force_bands {level: count}: test hook
    per_job = {}
    level_cache = {}
    for ji, job in enumerate(jobs):
        if job[6] != rank:
            continue
        lv = job[0]
        if lv not in level_cache:
            level_cache = {lv: pyr.level(levels[lv])}             # jobs are sorted by level: keep one level image alive
        per_job[ji] = run_band(model, level_cache[lv], job, templates, prob_thresh, rf, levels[lv])
    if per_job:
        counts = torch.cat([c for _b, _s, c in per_job.values()]).cpu().tolist()
        per_job = {ji: (b[:n], s[:n]) for (ji, (b, s, _c)), n in zip(per_job.items(), counts)}
    boxes, scores = gather_level_candidates(per_job, len(jobs), group=group, dst=0, device=device)
    if rank != 0:
        return (None, jobs) if return_plan else None
    dets = boxes[ops.nms_keep(boxes, scores, nms_thresh)].cpu().numpy()
    return (dets, jobs) if return_plan else dets
