"""tf_nms: size-class grid (algorithm 3) vs 1-D sweep (2) on the benchmark box sets: random boxes with ~25 % kept, the sparser
round-1 set, and dense pyramid candidates (inference leg)."""
import os, sys, json, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from tinyfaces_b200 import ops, synthetic, inference_bench
from tinyfaces_b200._lib import lib
dev = torch.device("cuda:0")
def run(tag, b, s):
    out = dict(set=tag, n=int(b.shape[0]))
    for algo in (2, 3):
        t = bench._event_time(lambda: ops.nms_device(b, s, 0.3, algo), 10)
        keep, cnt = ops.nms_device(b, s, 0.3, algo)
        out["ms_algo%d" % algo] = round(t * 1e3, 3); out["kept_algo%d" % algo] = int(cnt.item())
        lib().tf_debug_set(13, 1); ops.nms_device(b, s, 0.3, algo); st = ops.nms_sweep_stats(b.shape[0], 8, dev); lib().tf_debug_set(13, 0)
        out["pairs_algo%d" % algo] = st["pair_tests"]; out["edges_algo%d" % algo] = st["edges"]
        for stop, name in ((1, "sort"), (2, "+cand"), (3, "+resolve")):
            lib().tf_debug_set(14, stop)
            out["%s_algo%d" % (name, algo)] = round(bench._event_time(lambda: ops.nms_device(b, s, 0.3, algo), 10) * 1e3, 3)
        lib().tf_debug_set(14, 0)
    print(json.dumps(out), flush=True)
for n in (100000, 1000000):
    b, s = synthetic.boxes(n, seed=0, extent=0.35 * 40.0 * math.sqrt(n / 4.0)); run("random 25%% kept", b.to(dev), s.to(dev))
b, s = synthetic.boxes(100000, seed=0); run("random 61% kept (round-1 set)", b.to(dev), s.to(dev))
m = inference_bench.make_calibrated_model(dev)
r = inference_bench.run(m, base=1250, target_candidates=100000, reps=1)
print(json.dumps({k: r[k] for k in ("candidates", "kept", "nms_ms", "nms_ms_by_algorithm", "nms_stats") if k in r}), flush=True)
