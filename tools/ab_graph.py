"""A/B of the training step at a given shape: eager torch-SGD path vs flat (bucketed SGD) vs CUDA-graph replay."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
from tinyfaces_b200 import synthetic
from tinyfaces_b200.models.loss import DetectionCriterion
from tinyfaces_b200.models.model import DetectionModel
from tinyfaces_b200.optim import FlatSGD
from tinyfaces_b200.trainer import GraphedTrainStep, train_step, train_step_flat

def run(B, H, W, mode, steps=10, precision="fast"):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = DetectionModel(pretrained_weights=None, num_templates=25).to(dev).train()
    m.precision = precision
    crit = DetectionCriterion(25, sampler="device")
    x = synthetic.images(B, H, W).to(dev)
    cm, rm = synthetic.targets(B, (H + 7) // 8, (W + 7) // 8)
    cm, rm = cm.to(dev), rm.to(dev)
    if mode == "torch":
        opt = torch.optim.SGD(m.learnable_parameters(1e-7), momentum=0.9, weight_decay=5e-4, fused=True)
        fn = lambda: train_step(m, crit, opt, x, cm.clone(), rm)
    else:
        opt = FlatSGD(m, m.learnable_parameters(1e-7), momentum=0.9, weight_decay=5e-4)
        if mode == "flat":
            fn = lambda: train_step_flat(m, crit, opt, x, cm.clone(), rm)
        else:
            g = GraphedTrainStep(m, crit, opt, x, cm, rm)
            fn = lambda: g(x, cm, rm)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = fn()
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    torch.cuda.synchronize()
    return dict(B=B, H=H, W=W, mode=mode, precision=precision, ms=e0.elapsed_time(e1) / steps, host_enqueue_ms=host_ms, loss=float(loss))

if __name__ == "__main__":
    cases = [(8, 960, 1280), (4, 500, 500)]
    for (B, H, W) in cases:
        for mode in ("torch", "flat", "graph"):
            try:
                print(json.dumps(run(B, H, W, mode)), flush=True)
            except Exception as ex:
                print(json.dumps(dict(B=B, H=H, W=W, mode=mode, error=str(ex)[:400])), flush=True)
            torch.cuda.empty_cache()
    for mode in ("flat", "graph"):
        try:
            print(json.dumps(run(8, 960, 1280, mode, steps=5, precision="parity")), flush=True)
        except Exception as ex:
            print(json.dumps(dict(mode="parity-" + mode, error=str(ex)[:400])), flush=True)
        torch.cuda.empty_cache()
