"""Pyramid inference benchmark (BASELINE.json configs[2]).  python tools/bench_inference.py [target_N] [base]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
import torch

from tinyfaces_b200 import inference_bench

if __name__ == "__main__":
    if "--flags" in sys.argv:                      # --flags key:value,key:value -> tf_debug_set experiment switches
        i = sys.argv.index("--flags")
        from tinyfaces_b200 import _lib
        for kv in sys.argv[i + 1].split(","):
            k, v = kv.split(":")
            _lib.check(_lib.lib().tf_debug_set(int(k), int(v)), "tf_debug_set")
        del sys.argv[i:i + 2]
    target_n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    base = int(sys.argv[2]) if len(sys.argv) > 2 else 1250
    m = inference_bench.make_calibrated_model(torch.device('cuda:0'))
    print(json.dumps(inference_bench.run(m, base=base, target_candidates=target_n)))
