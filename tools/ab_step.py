"""A/B timing of the training step under the library's experiment switches (tf_debug_set keys), one subprocess
per configuration so that a trapped kernel cannot poison the others.  Results -> gpurun_out/ab_step.jsonl.

    python tools/ab_step.py [name=key:value,key:value ...]

Each record: ms/step (CUDA events over `steps` resident-input steps), the phase split (forward / loss / backward /
SGD, events on the caller's stream), the loss and a gradient checksum (configurations must agree on both).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
sys.path.insert(0, ROOT)

# keys: 3:2 = tail split-K off; 4:1 = 2-CTA GEMM; 5:1/2/3 = weight-gradient schedule 0/1/2; 6:1 = chain on the caller's stream
DEFAULT = [
    "legacy=5:1,6:1,3:2",      # round-1 behaviour: wgrad forked before its dgrad, no priority stream, no split-K
    "legacy_splitk=5:1,6:1",
    "default=",
    "default_2cta=4:1",
]


def run_case(spec, steps=8, warmup=3, B=8, H=960, W=1280):
    import torch
    from tinyfaces_b200 import _lib, ops, synthetic
    from tinyfaces_b200.models.loss import DetectionCriterion
    from tinyfaces_b200.models.model import DetectionModel
    name, _, flags = spec.partition("=")
    for kv in filter(None, flags.split(",")):
        if kv.startswith("env:"):                     # env:NAME:VALUE (read by the library at first use)
            _, k, v = kv.split(":")
            os.environ[k] = v
            continue
        k, v = kv.split(":")
        _lib.check(_lib.lib().tf_debug_set(int(k), int(v)), "tf_debug_set")
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = DetectionModel(pretrained_weights=None, num_templates=25).to(dev)
    model.train()
    crit = DetectionCriterion(25, sampler="device", seed=0)
    opt = torch.optim.SGD(model.learnable_parameters(1e-4), momentum=0.9, weight_decay=5e-4, fused=True)
    H3, W3 = (H + 7) // 8, (W + 7) // 8
    img = synthetic.images(B, H, W, seed=0).to(dev)
    cm, rm = synthetic.targets(B, H3, W3, 25, seed=0)
    cm, rm = cm.to(dev), rm.to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    phases = [0.0, 0.0, 0.0, 0.0]
    loss = None
    for it in range(warmup + steps):
        if it == warmup:
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
        ev[0].record()
        out = model(img)
        ev[1].record()
        loss = crit(out, cm.clone(), rm)
        ev[2].record()
        opt.zero_grad()
        loss.backward()
        ev[3].record()
        if it == 0:                            # identical weights in every configuration: these must agree
            named = dict(model.named_parameters())
            first = dict(loss=float(loss), **{k: float(named[k].grad.double().abs().sum()) for k in
                         ("model.conv1.weight", "model.layer1.0.conv2.weight", "model.layer2.3.conv1.weight",
                          "model.layer3.0.downsample.0.weight", "model.layer3.11.conv2.weight",
                          "model.layer3.22.conv3.weight", "model.layer3.22.bn3.weight", "score_res4.weight")})
        opt.step()
        ev[4].record()
        if it >= warmup:
            torch.cuda.synchronize()          # phase timing needs the events; costs one sync per step
            for k in range(4):
                phases[k] += ev[k].elapsed_time(ev[k + 1]) / steps
    t1.record()
    torch.cuda.synchronize()
    gsum = sum(float(p.grad.double().abs().sum()) for p in model.parameters() if p.grad is not None)
    # an un-instrumented timing loop (no per-step sync): the number bench.py reports
    for _ in range(2):
        loss = crit(model(img), cm.clone(), rm); opt.zero_grad(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(steps):
        loss = crit(model(img), cm.clone(), rm); opt.zero_grad(); loss.backward(); opt.step()
    t1.record()
    torch.cuda.synchronize()
    return dict(name=name, flags=flags, ms_per_step=t0.elapsed_time(t1) / steps, fwd_ms=phases[0], loss_ms=phases[1],
                bwd_ms=phases[2], sgd_ms=phases[3], loss=float(loss), grad_abs_sum=gsum, flag=ops.gemm_error_flag(),
                first_step=first)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        print("RESULT " + json.dumps(run_case(sys.argv[2])))
        sys.exit(0)
    specs = sys.argv[1:] or DEFAULT
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_step.jsonl"), "a") as f:
        for spec in specs:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", spec], capture_output=True,
                                   text=True, timeout=240)
                line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                rec = json.loads(line[0][7:]) if line else dict(name=spec, rc=r.returncode, err=r.stderr[-1500:])
            except subprocess.TimeoutExpired:
                rec = dict(name=spec, timeout=True)
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:1200])
