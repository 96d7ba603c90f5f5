"""Isolated timing of the BatchNorm kernels at the layer-3 shapes + the whole step (graph replay)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda:0")
pk = bench.peaks()
for r in bench.elementwise_classes(dev, pk):
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
st = bench.build_step(dev, 8, 960, 1280, "fast", 0, None)
ms = bench.timed_steps(st["step"], 10, 3, 1, dev) / 10
print(json.dumps(dict(step_ms=ms, graph=st["graphed"] is not None)), flush=True)
