"""Timing helper for the pyramid-inference workload (BASELINE.json configs[2]): per-level GPU pyramid + forward +
device decode, then one global NMS.  Used by bench.py and tools/bench_inference.py."""
import json
import os

import numpy as np
import torch
from torchvision import transforms

from . import ops
from .evaluation import _Pyramid, decode_level

_HERE = os.path.dirname(os.path.abspath(__file__))
RF = {"size": [859, 859], "stride": [8, 8], "offset": [-1, -1]}          # wider_face.py:55
FWD_GFLOP_1250 = {-2: 28.4, -1: 114.4, 0: 448.0, 1: 1773.3, 2: 7057.3}    # SURVEY.md section 8a (base 1250)


def load_templates():
    with open(os.path.join(_HERE, "templates.json")) as f:
        return np.array(json.load(f), dtype=np.float64)


def make_calibrated_model(device, seed=0):
    """Random-init detector whose eval-mode logits are finite: bn3 gamma = 0.25 and BN running
    statistics taken from one training-mode pass (momentum 1.0) over a random batch."""
    from .models.model import DetectionModel
    torch.manual_seed(seed)
    m = DetectionModel(pretrained_weights=None, num_templates=25)
    for name, p in m.named_parameters():
        if name.endswith("bn3.weight"):
            p.data.fill_(0.25)
    m = m.to(device)
    m.train()
    m.bn_momentum = 1.0
    with torch.no_grad():
        m(torch.randn(2, 3, 512, 512, device=device))
    m.bn_momentum = 0.1
    m.eval()
    return m


def threshold_for(model, img, img_transforms, scales, target_candidates, device):
    """Probability threshold that yields ~target_candidates over all levels (a random-init net has no meaningful scores)."""
    pyr = _Pyramid(img, img_transforms, device)
    probs = []
    with torch.no_grad():
        for s in [2 ** x for x in scales]:
            o = model(pyr.level(s))
            pr = torch.sigmoid(o[:, :25])
            pr[:, :, :, [0, 1, 2, 3] + list(range(12, 25))] = 0          # utils.py:44 column quirk
            probs.append(pr.flatten())
    allp = torch.cat(probs)
    return float(torch.topk(allp, min(target_candidates, allp.numel() - 1)).values[-1])


def run(model, base=1250, scales=(-2, -1, 0, 1, 2), target_candidates=100000, nms_thresh=0.3, seed=1, reps=2):
    """Returns a dict of per-stage device times (ms) for one synthetic base x base image."""
    dev = next(model.parameters()).device
    templates = load_templates()
    img = torch.rand(3, base, base, generator=torch.Generator().manual_seed(seed))
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    was_training = model.training
    model.eval()
    pyr = _Pyramid(img, tf, dev)
    lv = [2 ** s for s in scales]
    thr = threshold_for(model, img, tf, scales, target_candidates, dev)
    out = None
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        pyr_ms, fwd_ms, dec_ms, boxes, scores = [], [], [], [], []
        with torch.no_grad():
            for s in lv:
                ev[0].record()
                x = pyr.level(s)
                ev[1].record()
                o = model(x)
                ev[2].record()
                b, sc = decode_level(o, templates, thr, RF, s)
                ev[3].record()
                torch.cuda.synchronize()
                pyr_ms.append(ev[0].elapsed_time(ev[1]))
                fwd_ms.append(ev[1].elapsed_time(ev[2]))
                dec_ms.append(ev[2].elapsed_time(ev[3]))
                boxes.append(b)
                scores.append(sc)
        bx, sx = torch.cat(boxes), torch.cat(scores)
        ev[0].record()
        keep, cnt = ops.nms_device(bx, sx, nms_thresh)
        ev[1].record()
        torch.cuda.synchronize()
        nms_ms = ev[0].elapsed_time(ev[1])
        n, kept = int(bx.shape[0]), int(cnt.item())
        overflow = kept < 0
        if overflow:                                   # conflict list larger than the workspace: the exact bit-matrix path
            ev[0].record()
            keep = ops.nms_keep(bx, sx, nms_thresh, 1)
            ev[1].record()
            torch.cuda.synchronize()
            nms_ms += ev[0].elapsed_time(ev[1])
            kept = int(keep.numel())
        out = dict(workload="BASELINE.json configs[2]: %d-scale pyramid of a %dx%d image (levels %s px) + dense NMS"
                            % (len(lv), base, base, [int(base * s) for s in lv]),
                   candidates=n, kept=kept, prob_thresh=thr, pyramid_ms=pyr_ms, forward_ms=fwd_ms, decode_ms=dec_ms,
                   nms_ms=nms_ms, total_gpu_ms=sum(pyr_ms) + sum(fwd_ms) + sum(dec_ms) + nms_ms,
                   nms_boxes_per_s=n / (nms_ms / 1e3) if nms_ms > 0 else None, nms_edge_list_overflow=overflow)
        if not overflow:
            try:
                out["nms_stats"] = ops.nms_sweep_stats(n, 8, dev)
                out["nms_ms_by_algorithm"] = {}
                for algo, name in ((2, "1-D sweep"), (3, "size-class grid")):
                    ops.nms_device(bx, sx, nms_thresh, algo)
                    ev[0].record()
                    ops.nms_device(bx, sx, nms_thresh, algo)
                    ev[1].record()
                    torch.cuda.synchronize()
                    out["nms_ms_by_algorithm"][name] = ev[0].elapsed_time(ev[1])
            except Exception:  # noqa: BLE001
                pass
        if base == 1250:
            out["forward_tflops_per_level"] = [FWD_GFLOP_1250[s] / ms for s, ms in zip(scales, fwd_ms) if s in FWD_GFLOP_1250]
    # the launch-bound small levels again, replayed from CUDA graphs (DetectionModel.cuda_graphs)
    try:
        model.cuda_graphs = True
        gms = []
        with torch.no_grad():
            for s in lv:
                x = pyr.level(s)
                if x.shape[2] * x.shape[3] > model.cuda_graph_max_pixels:
                    gms.append(None)
                    continue
                model(x)                                     # capture
                model(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    model(x)
                e1.record()
                torch.cuda.synchronize()
                gms.append(e0.elapsed_time(e1) / 5)
        out["forward_ms_cuda_graph"] = gms
    except Exception as ex:  # noqa: BLE001
        out["forward_ms_cuda_graph"] = dict(error=str(ex)[:200])
    finally:
        model.cuda_graphs = False
    if was_training:
        model.train()
    return out
