#!/usr/bin/env python
"""Headline benchmark: detector training step (forward + loss + backward + SGD) images/s, plus NMS boxes/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1 = BASELINE.json configs[1]: batch-8 1280x960 synthetic images (tensor 8x3x960x1280), full
fwd+loss+bwd(+SGD) on one B200.  N>1 (torchrun, one rank per GPU): the same per-GPU batch on every rank with a
SUM gradient all-reduce over NCCL -> weak scaling.  `--impl reference` times the reference's CPU path (the oracle
port of it -- /root/reference does not exist on the GPU box) on the host cores.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

H_IMG, W_IMG, B_PER_GPU, T = 960, 1280, 8, 25
FWD_GFLOP_PER_IMG, STEP_GFLOP_PER_IMG = 346.1, 1032.5
METRIC = "train-step images/sec (fwd+loss+bwd+SGD), batch-8 1280x960 per GPU"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf32_burst=d["bf16_tflops"] / 2, tf32_sustained=d["bf16_tflops_sustained"] / 2,
                    source="MEASURED_PEAKS.json (bf16 cuBLAS / 2 for TF32)")
    return dict(hbm_gbs=6650.0, tf32_burst=795.0, tf32_sustained=700.0, source="fallback (B200_PROFILING.md) / 2 for TF32")


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(o[0]))
                self.max_mhz = float(o[1])
                for n, v in zip(names, o[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


# ------------------------------------------------------------------------------------------------ reference arm
def _best_thread_count():
    """torch CPU convs stop scaling (and then regress) well before 128 threads: pick the fastest count on a tiny conv."""
    import torch.nn.functional as F
    cores = os.cpu_count() or 1
    x = torch.randn(1, 256, 60, 80)            # the dominant layer3 3x3 at the sampled size
    w = torch.randn(256, 256, 3, 3)
    best, best_t = 1, 1e30
    for n in sorted({c for c in (8, 16, 32, 64, cores) if c <= cores}):
        torch.set_num_threads(n)
        F.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            F.conv2d(x, w, padding=1)
        t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = n, t
    return best


def cpu_reference_step_rate(steps, warmup, sample_hw=(480, 640), batch=1):
    """The reference's CPU path (oracle port: same torch CPU ops, same loss incl. the numpy sampler), fp32, on the
    host threads that run it fastest.  The sample is B=1 at 480x640 (a quarter of the 960x1280 image: per-image conv
    cost is proportional to area), scaled to 960x1280-image units.  Returns (images/s, threads, sample description)."""
    from oracle import loss_oracle, model_oracle, synth
    cores = _best_thread_count()
    torch.set_num_threads(cores)
    H, W = sample_hw
    sd = synth.synthetic_state_dict(seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running" not in k and not k.startswith("model.fc")}
    state = dict(sd)
    state.update(params)
    opt = torch.optim.SGD([p for k, p in params.items() if k != "score4_upsample.weight"], lr=1e-4, momentum=0.9,
                          weight_decay=5e-4)
    x = torch.randn(batch, 3, H, W, generator=torch.Generator().manual_seed(0))
    H3, W3 = (H + 7) // 8, (W + 7) // 8
    cm, rm = synth.synthetic_targets(batch, H3, W3, seed=0)
    np.random.seed(0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        new = {}
        out = model_oracle.forward(state, x, training=True, new_stats=new)
        res = loss_oracle.criterion(out.detach().numpy(), cm.copy(), rm)
        opt.zero_grad()
        out.backward(torch.from_numpy(res["grad"]))
        opt.step()
        state.update(new)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    area = (H * W) / float(H_IMG * W_IMG)
    return batch * area / t, cores, ("%d timed step(s) of B=%d %dx%d on %d of %d host threads (mean %.2f s/step), scaled by "
                                     "area to %dx%d images" % (steps, batch, H, W, cores, os.cpu_count() or 1, t, H_IMG, W_IMG))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    v, cores, sample = cpu_reference_step_rate(steps, warmup)
    line = dict(impl="reference", metric=METRIC, value=v, unit="images/s", n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=1000.0 / v, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32",
                data="synthetic", config=dict(workload="batch-8 1280x960 fwd+loss+bwd+SGD per GPU (CPU arm samples B=1)",
                                              per_gpu_batch=B_PER_GPU, image=[H_IMG, W_IMG]),
                cpu_baseline=dict(value=v, unit="images/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nms-n", type=int, default=100000)
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--breakdown", default="", help="write a per-kernel time table of one step to this file")
    ap.add_argument("--profile-mode", action="store_true", help="only the timed steps (for runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not args.profile_mode:
        args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from tinyfaces_b200 import ops, synthetic
    from tinyfaces_b200.models.loss import DetectionCriterion
    from tinyfaces_b200.models.model import DetectionModel
    from tinyfaces_b200.trainer import train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    torch.manual_seed(0)
    model = DetectionModel(pretrained_weights=None, num_templates=T).to(dev)
    model.train()
    crit = DetectionCriterion(T, sampler="device", seed=rank)
    # main.py:25-27 defaults (momentum .9, weight decay 5e-4) with the learning rate scaled down for this UNTRAINED net: the
    # reference's 1e-4 assumes ImageNet weights; on random-init weights and a summed loss it diverges to inf within a few
    # steps (the optimizer's work per step does not depend on the value)
    opt = torch.optim.SGD(model.learnable_parameters(1e-7), momentum=0.9, weight_decay=5e-4, fused=True)
    B = B_PER_GPU
    H3, W3 = (H_IMG + 7) // 8, (W_IMG + 7) // 8
    img_h = synthetic.images(B, H_IMG, W_IMG, seed=rank).pin_memory()
    cm_h, rm_h = synthetic.targets(B, H3, W3, T, seed=rank)
    cm_h, rm_h = cm_h.pin_memory(), rm_h.pin_memory()
    img_d, cm_d, rm_d = img_h.to(dev), cm_h.to(dev), rm_h.to(dev)

    def step_resident():
        return train_step(model, crit, opt, img_d, cm_d.clone(), rm_d)

    def run_e2e(steps):
        """`steps` steps through the public loop (trainer.train_pipelined): every step copies its inputs from pinned host
        memory (double-buffered on a copy stream) and its loss is read back to the host (one step late)."""
        from tinyfaces_b200.trainer import train_pipelined
        return list(train_pipelined(model, crit, opt, ((img_h, cm_h, rm_h) for _ in range(steps)), dev))

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(step_resident, args.steps, args.warmup)
    if args.profile_mode:
        sampler.stop_flag = True
        if rank == 0:
            print(json.dumps(dict(profile_mode=True, ms_per_step=ms_total / args.steps)))
        return
    run_e2e(2)                                                       # warm-up of the pipelined loop (slot allocation)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    e2e_losses = run_e2e(args.steps)
    ee1.record()
    torch.cuda.synchronize()
    ms_e2e_t = torch.tensor([ee0.elapsed_time(ee1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms_e2e_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e_t.item())
    assert len(e2e_losses) == args.steps
    sampler.stop_flag = True                     # clocks are sampled over both timed regions (resident + end-to-end)
    imgs = B * world * args.steps
    value = imgs / (ms_total / 1000.0)
    e2e_value = imgs / (ms_e2e / 1000.0)
    h2d = img_h.numel() * 4 + cm_h.numel() * 4 + rm_h.numel() * 4

    line = dict(metric=METRIC, value=value, unit="images/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="tf32",
                data="synthetic",
                config=dict(workload="BASELINE.json configs[1]: batch-8 1280x960 (8x3x960x1280) fwd+loss+bwd+SGD per GPU",
                            per_gpu_batch=B, global_batch=B * world, image=[H_IMG, W_IMG], templates=T,
                            precision="fast (1xTF32 operands, fp32 accumulate/storage)", sampler="device",
                            parallelism="dp%d (batch-sharded, SUM grad all-reduce, per-shard BN)" % world,
                            l2="step working set (~28 GB) >> 126 MB L2; no explicit flush"),
                e2e=dict(value=e2e_value, unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                         ms_per_step=ms_e2e / args.steps,
                         how="trainer.train_pipelined: pinned host batch -> device every step (double-buffered on a copy "
                             "stream, overlapping the previous step), loss read back to the host every step (one step late)"),
                clocks=sampler.summary())
    line["loss"] = dict(first=e2e_losses[0], last=e2e_losses[-1], finite=bool(np.all(np.isfinite(e2e_losses))))
    line["step_tflops"] = STEP_GFLOP_PER_IMG * value / 1000.0
    line["step_frac_of_tf32_sustained"] = line["step_tflops"] / (pk["tf32_sustained"] * world)

    # ---- isolated micro-benchmarks first (roofline kernel, NMS, pyramid inference): torch.profiler leaves CUPTI attached,
    #      which slows the host round trips of the latency-bound NMS afterwards
    if rank == 0:
        # ---- roofline of the dominant kernel: layer3 3x3 256->256 at this config (8x60x80), isolated, CUDA events
        Bc, Hc, Wc, C = B, H3 // 2, W3 // 2, 256
        xc = torch.randn(Bc, Hc, Wc, C, device=dev)
        wc = torch.randn(C, 9, C, device=dev) * 0.02
        yc = torch.empty(Bc, Hc, Wc, C, device=dev)
        for _ in range(5):
            ops.conv2d_nhwc(xc, wc, 3, out=yc)
        torch.cuda.synchronize()
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.conv2d_nhwc(xc, wc, 3, out=yc)
        e1.record()
        torch.cuda.synchronize()
        t_launch = e0.elapsed_time(e1) / reps / 1000.0
        flops = 2.0 * Bc * Hc * Wc * C * C * 9
        ach = flops / t_launch / 1e12
        traffic, pipe = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                rec = json.load(f)["conv_gemm_kernel<256> 3x3 256->256 M=38400"]
            traffic, pipe = rec["dram_read_bytes"] + rec["dram_write_bytes"], rec["tensor_pipe_active_pct"]
        except Exception:  # noqa: BLE001
            pass
        line["roofline"] = dict(bound="tensor", kernel="conv_gemm_kernel<256> (3x3 256->256, M=%d)" % (Bc * Hc * Wc),
                                achieved=ach, peak=pk["tf32_burst"], unit="TFLOP/s", frac=ach / pk["tf32_burst"],
                                traffic=traffic, traffic_source="profiles/roofline_traffic.json (ncu --set full)",
                                ncu_tensor_pipe_active_pct=pipe, algorithmic_flops_per_launch=flops,
                                peak_source=pk["source"], us_per_launch=t_launch * 1e6)

        # ---- NMS boxes/s (the second half of BASELINE.json's metric), N random boxes, float64
        bx, sc = synthetic.boxes(args.nms_n, seed=0)
        bx, sc = bx.to(dev), sc.to(dev)
        for _ in range(2):
            ops.nms_device(bx, sc, 0.3)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            keep, cnt = ops.nms_device(bx, sc, 0.3)
        e1.record()
        torch.cuda.synchronize()
        t_nms = e0.elapsed_time(e1) / 5 / 1000.0
        n, k = args.nms_n, int(cnt.item())
        nms_bytes = n * 40 + 2 * n * 8 + 2 * ((n + 63) // 64) * min(n, 32768) * 8 + k * 8
        line["nms"] = dict(n=n, kept=k, boxes_per_s=n / t_nms, ms=t_nms * 1e3, dtype="f64",
                           algorithmic_gbs=nms_bytes / t_nms / 1e9, hbm_frac=nms_bytes / t_nms / 1e9 / pk["hbm_gbs"],
                           note="pair-test (ALU) bound at this N, see DESIGN.md")

        # ---- pyramid inference (BASELINE.json configs[2])
        if not args.no_inference:
            try:
                from tinyfaces_b200 import inference_bench
                imodel = inference_bench.make_calibrated_model(dev)      # fresh weights with calibrated BN statistics
                line["inference"] = inference_bench.run(imodel, base=1250, target_candidates=args.nms_n)
                del imodel
            except Exception as ex:  # noqa: BLE001
                line["inference"] = dict(error=str(ex)[:300])

    # ---- kernel launch census of one step (CUPTI via torch.profiler; not timed).  The step contains the gradient
    #      all-reduce, so EVERY rank runs it; only rank 0 records.
    prof = None
    try:
        if rank == 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step_resident()
                torch.cuda.synchronize()
        else:
            step_resident()
            torch.cuda.synchronize()
    except Exception as ex:      # noqa: BLE001
        line["profiler_error"] = str(ex)[:200]
    if rank == 0:
        try:
            if prof is None:
                raise RuntimeError("no profile")
            ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
            mine = [e for e in ev if any(s in e.name for s in ("conv_gemm_kernel", "conv_wgrad_kernel", "kernel<", "_kernel"))
                    and "at::native" not in e.name and "nccl" not in e.name.lower()]
            line["gpu_launches"] = len(mine)
            line["gpu_launches_all"] = len(ev)
            tot = sum(e.device_time for e in ev) or 1.0
            gemm = sum(e.device_time for e in ev if "conv_gemm_kernel" in e.name or "conv_wgrad_kernel" in e.name)
            line["gemm_share_of_step"] = gemm / tot
            if args.breakdown:
                agg = {}
                for e in ev:
                    a = agg.setdefault(e.name[:110], [0, 0.0])
                    a[0] += 1
                    a[1] += e.device_time
                with open(args.breakdown, "w") as f:
                    f.write("# one training step, batch-8 960x1280, torch.profiler (CUPTI) device times\n")
                    f.write("%-112s %6s %10s %6s\n" % ("kernel", "calls", "total_us", "share"))
                    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                        f.write("%-112s %6d %10.1f %5.1f%%\n" % (k, n, t, 100.0 * t / tot))
        except Exception as ex:      # noqa: BLE001
            line["gpu_launches"] = None
            line["profiler_error"] = str(ex)[:200]

        # ---- CPU baseline beside it (bounded sample)
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, cores, sample = cpu_reference_step_rate(steps=2, warmup=1)
                line["cpu_baseline"] = dict(value=v, unit="images/s", cores=cores, kind="port", sample=sample)
                from oracle import nms_oracle
                nb, ns = synthetic.boxes(20000, seed=0)
                t0 = time.perf_counter()
                nms_oracle.nms(nb.numpy(), ns.numpy(), 0.3)
                line["cpu_baseline"]["nms_boxes_per_s_n20000"] = 20000 / (time.perf_counter() - t0)
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = dict(error=str(ex)[:200])
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
