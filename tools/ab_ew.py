"""Same-box A/B of the BatchNorm-backward apply kernel (tf_debug_set(15, 16) = per-channel vectors re-read per float4, the
round-1 form; default: hoisted into registers): isolated + whole step."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from tinyfaces_b200._lib import lib
dev = torch.device("cuda:0")
pk = bench.peaks()
for flag in (0, 32, 0, 32):
    lib().tf_debug_set(15, flag)
    for r in bench.elementwise_classes(dev, pk):
        if "backward" in r["name"]:
            print(json.dumps(dict(flag=flag, name=r["name"][:24], us=round(r["us"], 1), frac=round(r["frac_of_hbm_peak"], 3))), flush=True)
for flag in (0, 32, 0, 32):
    lib().tf_debug_set(15, flag)
    st = bench.build_step(dev, 8, 960, 1280, "fast", 0, None)
    ms = bench.timed_steps(st["step"], 10, 3, 1, dev) / 10
    print(json.dumps(dict(flag=flag, step_ms=round(ms, 3))), flush=True)
    del st; torch.cuda.empty_cache()
