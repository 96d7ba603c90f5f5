"""Pyramid inference benchmark (BASELINE.json configs[2]).  python tools/bench_inference.py [target_N] [base]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
import torch

from tinyfaces_b200 import inference_bench
from tinyfaces_b200.models.model import DetectionModel

if __name__ == "__main__":
    target_n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    base = int(sys.argv[2]) if len(sys.argv) > 2 else 1250
    torch.manual_seed(0)
    m = DetectionModel(pretrained_weights=None, num_templates=25)
    for name, p in m.named_parameters():
        if name.endswith("bn3.weight"):
            p.data.fill_(0.25)
    m = m.cuda()
    m.train()
    m.bn_momentum = 1.0                      # calibrate running statistics, else eval logits overflow
    with torch.no_grad():
        m(torch.randn(2, 3, 512, 512, device="cuda"))
    m.bn_momentum = 0.1
    print(json.dumps(inference_bench.run(m, base=base, target_candidates=target_n)))
