"""Where does get_detections spend its time on the 8-scale pyramid (BASELINE configs[4], one GPU)?  Per-phase host wall
clock with a synchronise after every phase (so the phases do not overlap as they do in the real call)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
import numpy as np
import torch
from torchvision import transforms
from tinyfaces_b200 import inference_bench, ops
from tinyfaces_b200.evaluation import _Pyramid, decode_level, get_detections

dev = torch.device("cuda:0")
m = inference_bench.make_calibrated_model(dev)
tpl = inference_bench.load_templates()
tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
img = torch.rand(3, 1250, 1250, generator=torch.Generator().manual_seed(1))
scales = (-1.5, -1, -0.5, 0, 0.5, 1, 1.5, 2)
thr = inference_bench.threshold_for(m, img, tf, scales, 100000, dev)

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for graphs in (False, True):
    m.cuda_graphs = graphs
    for rep in range(3):
        t0 = T()
        with torch.no_grad():
            d = get_detections(m, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev)
        t1 = T()
    print("cuda_graphs=%s: get_detections %.1f ms (%d dets)" % (graphs, (t1 - t0) * 1e3, len(d)), flush=True)
m.cuda_graphs = False
with torch.no_grad():
    t = T(); pyr = _Pyramid(img, tf, dev); print("pyramid init %.2f ms" % ((T() - t) * 1e3))
    parts = []
    for s in scales:
        sc = 2 ** s
        t = T(); x = pyr.level(sc); t_l = T() - t
        t = time.perf_counter(); o = m(x); t_host = time.perf_counter() - t; t_f = T() - t
        t = T(); p = decode_level(o, tpl, thr, inference_bench.RF, sc, sync=False); t_d = T() - t
        parts.append(p)
        print("level %5d px: pyramid %.2f ms, forward %.2f ms (host enqueue %.2f ms), decode %.2f ms" % (x.shape[2], t_l * 1e3, t_f * 1e3, t_host * 1e3, t_d * 1e3), flush=True)
    t = T(); counts = torch.cat([c for _b, _s, c in parts]).cpu().tolist(); print("counts %.2f ms" % ((T() - t) * 1e3), counts)
    t = T()
    boxes = torch.cat([b[:n] for (b, _s, _c), n in zip(parts, counts)]); scores = torch.cat([s[:n] for (_b, s, _c), n in zip(parts, counts)])
    print("cat %.2f ms" % ((T() - t) * 1e3))
    t = T(); keep, cnt = ops.nms_device(boxes, scores, 0.3); k = int(cnt.item()); print("nms %.2f ms count %d" % ((T() - t) * 1e3, k), ops.nms_sweep_stats(boxes.shape[0], 8, dev))
    t = T(); dd = boxes[keep[:max(k, 0)]].cpu().numpy(); print("gather+d2h %.2f ms" % ((T() - t) * 1e3))
