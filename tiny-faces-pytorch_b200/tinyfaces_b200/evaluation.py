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


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms(boxes[N,4], scores[N], iou_threshold) -> int64[K] (descending score), bit-identical
    keep indices.  CPU inputs are moved to the current CUDA device and the result is returned on the input's
    device, like the op it replaces."""
    if boxes.dtype != scores.dtype:
        raise RuntimeError("nms: boxes and scores should have the same dtype")
    src = boxes.device
    dev = src if boxes.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return ops.nms_keep(boxes.to(dev), scores.to(dev), iou_threshold).to(src)


def decode_level(output, templates, prob_thresh, rf, scale, bug_compat=True, sync=True):
    """Device-side replacement of evaluation.py:61-71 for one pyramid level.
    output: [B,5T,H,W] CUDA tensor.  Returns (boxes f64 [N,4], scores f64 [N]) on the device; with sync=False the
    capacity-sized buffers and the device-side count are returned instead (boxes, scores, count) so that the caller can
    read all counts with ONE host synchronisation after every level has been enqueued."""
    B, C, H, W = output.shape
    T = templates.shape[0]
    if bug_compat and W < 25:
        raise IndexError("index 24 is out of bounds for axis 2 with size %d" % W)   # the reference's failure mode
    inv = _bitmask(invalid_template_ids(templates, scale))
    hw = H * W
    strides = (C * hw, W, 1, hw)                       # (b, y, x, c) element strides of an NCHW tensor
    out = output.contiguous()
    boxes, scores, _, count = ops.decode_device(out, out[:, T:], None, strides, strides, B, H, W, T, templates,
                                                prob_thresh, inv if bug_compat else 0, 0 if bug_compat else inv, rf, scale)
    if not sync:
        return boxes, scores, count
    n = int(count.item())
    return boxes[:n], scores[:n]


def get_detections(model, img, templates, rf, img_transforms, prob_thresh=0.65, nms_thresh=0.3, scales=(-2, -1, 0, 1),
                   device=None, return_scores=False, gpu_pyramid=True):
    """evaluation.py:20-87.  img: CHW float tensor in [0,1]; scales are exponents of 2; returns ndarray [K,4] float64."""
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    model = model.to(device)
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
    total = torch.stack(all_counts).sum(0)                       # per-level totals (each level lives on one rank)
    mine = sorted(per_level)
    payload = torch.cat([torch.cat([per_level[lv][0], per_level[lv][1][:, None]], dim=1) for lv in mine]) \
        if mine else torch.zeros((0, 5), dtype=torch.float64, device=dev)
    sizes = [int(c.sum().item()) for c in all_counts]
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
            n = int(all_counts[r][lv].item())
            if n:
                chunks[lv] = gathered[r][off:off + n]
                off += n
    cat = torch.cat([c for c in chunks if c is not None]) if any(c is not None for c in chunks) \
        else torch.zeros((0, 5), dtype=torch.float64, device=dev)
    assert cat.shape[0] == int(total.sum().item())
    return cat[:, :4].contiguous(), cat[:, 4].contiguous()


def get_detections_sharded(model, img, templates, rf, img_transforms, prob_thresh=0.65, nms_thresh=0.3,
                           scales=(-2, -1, 0, 1), device=None, group=None):
    """get_detections with one pyramid level per rank (round-robin when there are more levels than ranks), a
    candidate gather to rank 0 and the global NMS there.  Rank 0 returns ndarray [K,4]; other ranks return None."""
    import torch.distributed as dist
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    model = model.to(device)
    model.eval()
    templates = np.asarray(templates, dtype=np.float64)
    pyr = _Pyramid(img, img_transforms, device)
    levels = [2 ** x for x in scales]
    # largest levels first onto distinct ranks (cost ~ 4^exponent): simple longest-processing-time packing
    order = sorted(range(len(levels)), key=lambda i: -levels[i])
    load = [0.0] * world
    owner = {}
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        owner[i] = r
        load[r] += levels[i] ** 2
    per_level = {}
    for i, scale in enumerate(levels):
        if owner[i] != rank:
            continue
        x = pyr.level(scale)
        with torch.no_grad():
            output = model(x)
        per_level[i] = decode_level(output, templates, prob_thresh, rf, scale)
    boxes, scores = gather_level_candidates(per_level, len(levels), group=group, dst=0, device=device)
    if rank != 0:
        return None
    return boxes[ops.nms_keep(boxes, scores, nms_thresh)].cpu().numpy()
