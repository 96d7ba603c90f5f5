"""Kernel timeline of ONE training step (torch.profiler / CUPTI): gpurun_out/timeline.csv with one row per kernel
(start us, duration us, stream, name) plus a summary of busy time per stream and the idle gaps on the chain."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile
    from tinyfaces_b200 import synthetic
    from tinyfaces_b200.models.loss import DetectionCriterion
    from tinyfaces_b200.models.model import DetectionModel
    from tinyfaces_b200.trainer import train_step
    dev = torch.device("cuda:0")
    B, H, W = 8, 960, 1280
    torch.manual_seed(0)
    model = DetectionModel(pretrained_weights=None, num_templates=25).to(dev)
    model.train()
    crit = DetectionCriterion(25, sampler="device", seed=0)
    opt = torch.optim.SGD(model.learnable_parameters(1e-4), momentum=0.9, weight_decay=5e-4, fused=True)
    img = synthetic.images(B, H, W, seed=0).to(dev)
    cm, rm = synthetic.targets(B, (H + 7) // 8, (W + 7) // 8, 25, seed=0)
    cm, rm = cm.to(dev), rm.to(dev)
    for _ in range(3):
        train_step(model, crit, opt, img, cm.clone(), rm)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):                       # the 2nd step is launched while the 1st still runs (steady state)
            train_step(model, crit, opt, img, cm.clone(), rm)
        torch.cuda.synchronize()
    out = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "timeline.csv")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tmp = out + ".json"
    prof.export_chrome_trace(tmp)
    import json
    with open(tmp) as f:
        tr = json.load(f)
    os.remove(tmp)
    ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    with open(out, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for e in ev:
            name = e["name"].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60].replace(",", ";")
            f.write("%.2f,%.2f,%s,%s\n" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"), name))
    print("wrote", out, len(ev), "events, span %.1f us" % (ev[-1]["ts"] + ev[-1]["dur"] - t0))


if __name__ == "__main__":
    main()
