#!/usr/bin/env python
"""Headline benchmark: detector training step (forward + loss + backward + all-reduce + SGD) images/s, plus NMS boxes/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1 = BASELINE.json configs[1]: batch-8 1280x960 synthetic images (tensor 8x3x960x1280), full
fwd+loss+bwd+SGD on one B200, the whole step replayed from ONE CUDA graph.  N>1 (torchrun, one rank per GPU): the same
per-GPU batch on every rank with the bucketed SUM gradient all-reduce over NCCL overlapped with the backward -> weak
scaling.  Extra keys on the same JSON line: `parity_mode` (the 3xTF32 arithmetic that meets the 1e-3 tolerance, timed, with
both modes' measured error against the reference's golden output at this very shape), `roofline` (the kernel class with
the largest time share) + `roofline_classes`, `nms` (stage breakdown, pair tests/s), `inference` (configs[2]),
`cfg3` (configs[3]: batch-32 500x500 strong scaling) and `cfg4` (configs[4]: 8-scale sharded pyramid inference).
`--impl reference` times the reference's CPU path (the oracle port of it -- /root/reference does not exist on the GPU
box) on the host cores.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import math
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
STEP_GFLOP_PER_IMG_500 = 218.0
METRIC = "train-step images/sec (fwd+loss+bwd+SGD), batch-8 1280x960 per GPU"
NMS_EXTENT_FACTOR = 0.35                                        # synthetic box density: ~25 % of the boxes survive (10-30 %)


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
    host threads that run it fastest.  The sample is B=1 at 480x640 (a QUARTER of the 960x1280 image area: per-image conv
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
    return batch * area / t, cores, ("%d timed step(s) of B=%d %dx%d (a quarter-area sample) on %d of %d host threads (mean %.2f s/step), "
                                     "scaled by area to %dx%d images" % (steps, batch, H, W, cores, os.cpu_count() or 1, t, H_IMG, W_IMG))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    v, cores, sample = cpu_reference_step_rate(steps, warmup)
    line = dict(impl="reference", metric=METRIC, value=v, unit="images/s", n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=1000.0 / v, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32",
                data="synthetic", config=dict(workload="batch-8 1280x960 fwd+loss+bwd+SGD per GPU (CPU arm samples B=1 at 480x640, quarter area)",
                                              per_gpu_batch=B_PER_GPU, image=[H_IMG, W_IMG]),
                cpu_baseline=dict(value=v, unit="images/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ helpers (GPU arm)
def _kernel_key(name):
    """'void (anonymous namespace)::conv_gemm_kernel<256, false, true, false>(...)' -> 'conv_gemm_kernel<256, false, true, false>'"""
    import re
    m = re.search(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^()]*>)?)\s*\(", name.replace("(anonymous namespace)::", ""))
    return (m.group(1) if m else name)[:80]



def _event_time(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1000.0


def build_step(dev, B, H, W, precision, seed, group, lr=1e-7, graph=True):
    """(model, criterion, optimizer, step callable, static device batch, graphed?)"""
    from tinyfaces_b200 import synthetic
    from tinyfaces_b200.models.loss import DetectionCriterion
    from tinyfaces_b200.models.model import DetectionModel
    from tinyfaces_b200.optim import FlatSGD
    from tinyfaces_b200.trainer import GraphedTrainStep, train_step_flat
    torch.manual_seed(0)
    model = DetectionModel(pretrained_weights=None, num_templates=T).to(dev)
    model.train()
    model.precision = precision
    crit = DetectionCriterion(T, sampler="device", seed=seed)
    # main.py:25-27 defaults (momentum .9, weight decay 5e-4) with the learning rate scaled down for this UNTRAINED net: the
    # reference's 1e-4 assumes ImageNet weights; on random-init weights and a summed loss it diverges to inf within a few
    # steps (the optimizer's work per step does not depend on the value)
    opt = FlatSGD(model, model.learnable_parameters(lr), momentum=0.9, weight_decay=5e-4)
    H3, W3 = (H + 7) // 8, (W + 7) // 8
    img_h = synthetic.images(B, H, W, seed=seed).pin_memory()
    cm_h, rm_h = synthetic.targets(B, H3, W3, T, seed=seed)
    cm_h, rm_h = cm_h.pin_memory(), rm_h.pin_memory()
    dev_batch = (img_h.to(dev), cm_h.to(dev), rm_h.to(dev))
    graphed, err = None, None
    if graph:
        try:
            graphed = GraphedTrainStep(model, crit, opt, *dev_batch, group=group, warmup=2)
        except Exception as ex:  # noqa: BLE001
            err = str(ex)[:300]
            graphed = None
            torch.cuda.synchronize()
    if graphed is not None:
        def step():
            return graphed(*dev_batch)
    else:
        def step():
            return train_step_flat(model, crit, opt, dev_batch[0], dev_batch[1].clone(), dev_batch[2], group)
    return dict(model=model, crit=crit, opt=opt, step=step, host=(img_h, cm_h, rm_h), dev=dev_batch, graphed=graphed, graph_error=err)


def timed_steps(fn, steps, warmup, world, dev):
    import torch.distributed as dist
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


def golden_forward_error(dev, precision):
    """Forward of the benchmark shape itself against the REFERENCE's output (tests/golden/cfg2_b8_fwd.npz, produced by
    oracle/make_golden.py from /root/reference): max-norm and L2 relative error of the 8x125x120x160 score map."""
    from tinyfaces_b200 import synthetic
    from tinyfaces_b200.models.model import DetectionModel
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg2_b8_fwd.npz"))
    m = DetectionModel(pretrained_weights=None, num_templates=T)
    m.load_state_dict(synthetic.state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1))
    m.precision = precision
    m = m.to(dev).train()
    B, C, H, W = (int(v) for v in g["shape"])
    x = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(int(g["seed"])))
    with torch.no_grad():
        out = m(x.to(dev))
    step = int(g["step"])
    sub = out[:, :, ::step, ::step].double().cpu().numpy()
    ref = g["out_sub"].astype(np.float64)
    res = dict(max_rel=float(np.abs(sub - ref).max() / float(g["out_max"])), l2_rel=float(np.linalg.norm(sub - ref) / np.linalg.norm(ref)))
    del m, out
    torch.cuda.empty_cache()
    return res


def roofline_classes(dev, pk):
    """The layer-3 GEMM classes of the step at M = 8*60*80 = 38400 pixels, each timed ALONE with CUDA events on the launch
    stream (-> compared with the burst peak), with their launch counts per training step."""
    from tinyfaces_b200 import ops
    B, Hc, Wc = B_PER_GPU, H_IMG // 16, W_IMG // 16
    rows = []

    def add(name, kernel, count, flops, fn):
        t = _event_time(fn, 40)
        rows.append(dict(name=name, kernel=kernel, launches_per_step=count, us_per_launch=t * 1e6, algorithmic_gflop=flops / 1e9,
                         tflops=flops / t / 1e12, frac_of_tf32_burst=flops / t / 1e12 / pk["tf32_burst"],
                         ms_per_step=count * t * 1e3))
    M = B * Hc * Wc
    for name, cin, cout, k, n_f, n_w in (("3x3 256->256", 256, 256, 3, 22 + 22, 22), ("1x1 256->1024", 256, 1024, 1, 23 + 22, 22 + 23),
                                          ("1x1 1024->256", 1024, 256, 1, 22 + 23, 22 + 23)):
        x = torch.randn(B, Hc, Wc, cin, device=dev)
        w = torch.randn(cout, k * k, cin, device=dev) * 0.02
        y = torch.empty(B, Hc, Wc, cout, device=dev)
        dy = torch.randn(B, Hc, Wc, cout, device=dev)
        dw = torch.zeros(cout, k * k, cin, device=dev)
        flops = 2.0 * M * cin * cout * k * k
        add("fprop/dgrad " + name, "conv_gemm_kernel / conv_gemm2_kernel", n_f, flops, lambda: ops.conv2d_nhwc(x, w, k, out=y))
        add("wgrad " + name, "conv_wgrad_kernel<256>", n_w, flops, lambda: ops.conv2d_wgrad_nhwc(x, dy, k, out=dw))
        del x, w, y, dy, dw
    return rows


def elementwise_classes(dev, pk):
    """The HBM-bound BatchNorm kernels at the layer-3 shapes (M = 38400 pixels, C = 1024 and 256), timed alone through their
    own C-ABI entry points with CUDA events.  Algorithmic bytes per element (fp32 tensors, 1-bit ReLU masks):
      BN backward  = colreduce_kernel<1> (dout 4 + y 4 + mask 1/8) + bn_bwd_apply_kernel (dout 4 + y 4 + mask 1/8 + dy 4) = 20.25 B
      BN forward   = colreduce_kernel<0>-free path: bn_apply_kernel with shortcut (y 4 + res 4 + out 4 + mask 1/8) = 12.125 B
                     (the statistics come out of the GEMM epilogue), without shortcut 8.125 B."""
    from tinyfaces_b200 import ops
    M = B_PER_GPU * (H_IMG // 16) * (W_IMG // 16)
    rows = []
    for C, n_bwd, n_fwd, with_res in ((1024, 23, 23, True), (256, 46, 46, False)):
        y = torch.randn(M, C, device=dev)
        gamma = torch.rand(C, device=dev) + 0.5
        beta = torch.randn(C, device=dev) * 0.1
        res = torch.randn(M, C, device=dev) if with_res else None
        out, mask, mean, rstd = ops.bn_train_fwd(y, gamma, beta, None, None, res, relu=True, round_tf32=True)
        dout = torch.randn(M, C, device=dev)
        t_b = _event_time(lambda: ops.bn_bwd(dout, mask, y, mean, rstd, gamma, round_tf32=True), 30)
        t_f = _event_time(lambda: ops.bn_train_fwd(y, gamma, beta, None, None, res, relu=True, round_tf32=True), 30)
        n = float(M) * C
        bb = 20.25 * n
        # the standalone forward entry point also runs the column-statistics pass the executor gets from the GEMM epilogue
        bf = (12.125 if with_res else 8.125) * n + 4.0 * n
        rows.append(dict(name="BN backward C=%d (colreduce_kernel<1> + bn_bwd_finalize + bn_bwd_apply_kernel)" % C, launches_per_step=n_bwd,
                         us=t_b * 1e6, algorithmic_mb=bb / 1e6, gbs=bb / t_b / 1e9, frac_of_hbm_peak=bb / t_b / 1e9 / pk["hbm_gbs"],
                         ms_per_step=n_bwd * t_b * 1e3))
        rows.append(dict(name="BN forward C=%d (column statistics + finalize + bn_apply_kernel%s)" % (C, ", shortcut add" if with_res else ""),
                         launches_per_step=n_fwd, us=t_f * 1e6, algorithmic_mb=bf / 1e6, gbs=bf / t_f / 1e9,
                         frac_of_hbm_peak=bf / t_f / 1e9 / pk["hbm_gbs"], ms_per_step=n_fwd * t_f * 1e3))
        del y, res, out, mask, dout
        torch.cuda.empty_cache()
    return rows


def nms_by_n(dev, sizes=(1000, 10000, 100000, 1000000)):
    """SURVEY 8d's NMS-only sweep: the synthetic float64 set of nms_report at N = 1e3 .. 1e6 (same density: ~25 % kept)."""
    from tinyfaces_b200 import ops, synthetic
    rows = []
    for n in sizes:
        bx, sc = synthetic.boxes(n, seed=0, extent=NMS_EXTENT_FACTOR * 40.0 * math.sqrt(n / 4.0))
        bxd, scd = bx.to(dev), sc.to(dev)
        t = _event_time(lambda: ops.nms_device(bxd, scd, 0.3), 10)
        _, cnt = ops.nms_device(bxd, scd, 0.3)
        k = int(cnt.item())
        rows.append(dict(n=n, ms=t * 1e3, boxes_per_s=n / t, kept_frac=k / n if k >= 0 else None,
                         candidates="bit-matrix blocks" if n < 4096 else ("1-D sweep" if n <= 300000 else "size-class grid")))
        del bxd, scd
    return rows


def nms_report(dev, n, pk, with_cpu):
    """tf_nms on n synthetic float64 boxes (centres U(0,S)^2 with S chosen so that ~25 % survive, sizes U(10,70)^2, 1 %
    exact duplicates): end-to-end time, stage breakdown (the library stops after a stage under tf_debug_set(14, k)),
    IoU pair tests and the bytes the algorithm that ran actually moves."""
    from tinyfaces_b200 import ops, synthetic
    from tinyfaces_b200._lib import lib
    bx, sc = synthetic.boxes(n, seed=0, extent=NMS_EXTENT_FACTOR * 40.0 * math.sqrt(n / 4.0))
    bxd, scd = bx.to(dev), sc.to(dev)
    t_all = _event_time(lambda: ops.nms_device(bxd, scd, 0.3), 10)
    t_grid = _event_time(lambda: ops.nms_device(bxd, scd, 0.3, 3), 10)             # candidate generation by the size-class grid instead
    keep, cnt = ops.nms_device(bxd, scd, 0.3)
    k = int(cnt.item())
    stages = {}
    for stop, name in ((1, "sort+gather"), (2, "+sweep"), (3, "+resolve")):
        lib().tf_debug_set(14, stop)
        stages[name] = _event_time(lambda: ops.nms_device(bxd, scd, 0.3), 10)
    lib().tf_debug_set(14, 0)
    lib().tf_debug_set(13, 1)
    ops.nms_device(bxd, scd, 0.3)
    st = ops.nms_sweep_stats(n, 8, dev)
    lib().tf_debug_set(13, 0)
    # bytes of the sort-and-sweep algorithm: boxes+scores in, two key/index sorts (64-bit and 32-bit keys, w+r per radix pass
    # is library-internal: counted once), rank- and x-ordered box copies, the edge list written once and read ~once, keep out
    nbytes = n * 40 + n * (8 + 4) * 2 + n * (4 + 4) * 2 + 2 * n * 40 * 2 + st["edges"] * 8 * 2 + n * 3 + k * 8
    rep = dict(n=n, kept=k, kept_frac=k / n, dtype="f64", thr=0.3, ms=t_all * 1e3, boxes_per_s=n / t_all,
               stage_ms={"sort+gather": stages["sort+gather"] * 1e3, "sweep": (stages["+sweep"] - stages["sort+gather"]) * 1e3,
                         "resolve": (stages["+resolve"] - stages["+sweep"]) * 1e3, "select": (t_all - stages["+resolve"]) * 1e3},
               conflict_edges=st["edges"], resolve_rounds=st["rounds"], iou_pair_tests=st["pair_tests"],
               pair_tests_per_s=st["pair_tests"] / t_all, all_pairs=n * (n - 1) // 2,
               algorithm="sort-and-sweep candidates (size-class grid above 3e5 boxes) + parallel fixed-point resolution; "
                         "stream-ordered, no host sync",
               ms_with_grid_candidates=t_grid * 1e3,
               algorithmic_bytes=nbytes, algorithmic_gbs=nbytes / t_all / 1e9, hbm_frac=nbytes / t_all / 1e9 / pk["hbm_gbs"],
               bound="latency / pair tests: %d dependent launches move %.0f MB -- the HBM roofline is not the limiter above N ~ 1e4 "
                     "(SURVEY 8d); boxes/s and pair tests/s are the honest figures" % (20, nbytes / 1e6))
    if with_cpu:
        try:
            from oracle import nms_oracle
            t0 = time.perf_counter()
            kc = nms_oracle.nms(bx.numpy(), sc.numpy(), 0.3)
            dt = time.perf_counter() - t0
            rep["cpu"] = dict(boxes_per_s=n / dt, s=dt, n=n, kept=int(len(kc)), same_keep=bool(np.array_equal(kc, keep[:k].cpu().numpy())),
                              kind="port (plain-C restatement of torchvision's CPU nms, 1 thread)")
        except Exception as ex:  # noqa: BLE001
            rep["cpu"] = dict(error=str(ex)[:200])
    return rep


def cfg3_report(dev, world, rank, group, steps):
    """BASELINE configs[3]: batch-32 500x500 training step, data-parallel: STRONG scaling (32 / N images per GPU)."""
    B = 32 // world
    st = build_step(dev, B, 500, 500, "fast", rank, group)
    ms = timed_steps(st["step"], steps, 3, world, dev)
    rep = dict(workload="BASELINE.json configs[3]: batch-32 500x500 step, %d image(s) per GPU, bucketed SUM all-reduce overlapped with "
                        "the backward + SGD as its epilogue" % B, scaling="strong", global_batch=32, per_gpu_batch=B, n_gpus=world,
               ms_per_step=ms / steps, images_per_s=32 * steps / (ms / 1e3), cuda_graph=st["graphed"] is not None,
               graph_error=st["graph_error"], buckets=len(st["opt"].flat.buckets),
               step_tflops_per_gpu=STEP_GFLOP_PER_IMG_500 * B / (ms / steps))
    del st
    torch.cuda.empty_cache()
    return rep


def cfg4_report(dev, world, rank, group, target_candidates=100000):
    """BASELINE configs[4]: 8-scale pyramid of a 1250x1250 image (441 ... 5000 px), sharded over the ranks by level AND by
    horizontal bands of the large levels (halo 448 px), candidate gather, global NMS on rank 0."""
    import torch.distributed as dist
    from torchvision import transforms
    from tinyfaces_b200 import inference_bench
    from tinyfaces_b200.evaluation import get_detections, get_detections_sharded
    model = inference_bench.make_calibrated_model(dev)
    tpl = inference_bench.load_templates()
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    img = torch.rand(3, 1250, 1250, generator=torch.Generator().manual_seed(1))
    scales = (-1.5, -1, -0.5, 0, 0.5, 1, 1.5, 2)
    thr = inference_bench.threshold_for(model, img, tf, scales, target_candidates, dev)
    rep = dict(workload="BASELINE.json configs[4]: 8-scale pyramid of a 1250x1250 image (levels %s px) + global NMS"
                        % [int(1250 * 2 ** s) for s in scales], n_gpus=world, prob_thresh=thr)

    def timed(fn):
        """device time of fn() on this rank's stream (CUDA events), MAX over the ranks"""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return r, float(ms.item())
    with torch.no_grad():
        single = None
        if rank == 0:
            get_detections(model, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev)      # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            single = get_detections(model, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev)
            e1.record()
            torch.cuda.synchronize()
            rep["single_gpu_ms"] = e0.elapsed_time(e1)
            rep["detections"] = int(len(single))
        if world > 1:
            for spatial, key in ((False, "level_sharded"), (True, "level+band_sharded")):
                get_detections_sharded(model, img, tpl, inference_bench.RF, tf, prob_thresh=thr, scales=scales, device=dev,
                                       group=group, spatial=spatial)                                                  # warm-up
                (res, ms) = timed(lambda: get_detections_sharded(model, img, tpl, inference_bench.RF, tf, prob_thresh=thr,
                                                                scales=scales, device=dev, group=group, spatial=spatial,
                                                                return_plan=True))
                if rank == 0:
                    dets, jobs = res
                    rep[key] = dict(ms=ms, speedup_vs_single=rep["single_gpu_ms"] / ms, identical_to_single_gpu=bool(np.array_equal(dets, single)),
                                    jobs=len(jobs), bands_per_level=[sum(1 for j in jobs if j[0] == lv) for lv in range(len(scales))])
    del model
    torch.cuda.empty_cache()
    return rep


# ------------------------------------------------------------------------------------------------ GPU arm
_PROGRESS = dict(line=None, phase="start", rank=0)


def _watchdog(deadline_s):
    """A stuck collective (or anything else) in one of the EXTRA sections must not cost the whole run: when the deadline
    passes, rank 0 prints the line as far as it got (with the phase it was stuck in) and every rank exits."""
    def run():
        time.sleep(deadline_s)
        if _PROGRESS["rank"] == 0 and _PROGRESS["line"] is not None:
            line = dict(_PROGRESS["line"])
            line["watchdog"] = "deadline of %d s passed during phase '%s': later sections are missing" % (deadline_s, _PROGRESS["phase"])
            try:
                print(json.dumps(line), flush=True)
            except Exception:  # noqa: BLE001
                pass
        os._exit(0)
    t = threading.Thread(target=run, daemon=True)
    t.start()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nms-n", type=int, default=100000)
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e only")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA-graph replay")
    ap.add_argument("--breakdown", default="", help="write a per-kernel time table of one step to this file")
    ap.add_argument("--profile-mode", action="store_true", help="only the timed steps, eager (for runs under ncu)")
    ap.add_argument("--deadline", type=int, default=780, help="seconds after which the line collected so far is printed and the run ends")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not args.profile_mode:
        args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from tinyfaces_b200.trainer import train_pipelined

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    _PROGRESS["rank"] = rank
    if not args.profile_mode:
        _watchdog(args.deadline)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    pk = peaks()

    st = build_step(dev, B_PER_GPU, H_IMG, W_IMG, "fast", rank, group, graph=not (args.no_graph or args.profile_mode))
    model, crit, opt, step_resident = st["model"], st["crit"], st["opt"], st["step"]
    img_h, cm_h, rm_h = st["host"]
    B = B_PER_GPU

    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed_steps(step_resident, args.steps, args.warmup, world, dev)
    if args.profile_mode:
        sampler.stop_flag = True
        if rank == 0:
            print(json.dumps(dict(profile_mode=True, ms_per_step=ms_total / args.steps)))
        return

    def run_e2e(steps):
        """`steps` steps through the public loop (trainer.train_pipelined): every step copies its inputs from pinned host
        memory (double-buffered on a copy stream) and its loss is read back to the host (one step late)."""
        return list(train_pipelined(model, crit, opt, ((img_h, cm_h, rm_h) for _ in range(steps)), dev, group=group,
                                    graph=st["graphed"] if st["graphed"] is not None else False))
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
                            precision="fast (1xTF32 operands, fp32 accumulate/storage); see parity_mode for the 3xTF32 arithmetic",
                            sampler="device", cuda_graph=st["graphed"] is not None, graph_error=st["graph_error"],
                            optimizer="library SGD kernel per gradient bucket (momentum .9, wd 5e-4), %d buckets" % len(opt.flat.buckets),
                            parallelism="dp%d (batch-sharded, bucketed SUM grad all-reduce overlapped with the backward, per-shard BN)" % world,
                            l2="step working set (~28 GB) >> 126 MB L2; no explicit flush"),
                e2e=dict(value=e2e_value, unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=12,
                         ms_per_step=ms_e2e / args.steps,
                         how="trainer.train_pipelined(graph=...): pinned host batch -> device every step (double-buffered on a copy "
                             "stream, overlapping the previous step), device->device into the graph's static inputs, graph replay, "
                             "loss read back to the host every step (one step late)"),
                clocks=sampler.summary())
    _PROGRESS["line"] = line
    _PROGRESS["phase"] = "launch census"
    line["loss"] = dict(first=e2e_losses[0], last=e2e_losses[-1], finite=bool(np.all(np.isfinite(e2e_losses))))
    line["step_tflops"] = STEP_GFLOP_PER_IMG * value / 1000.0
    line["step_frac_of_tf32_sustained"] = line["step_tflops"] / (pk["tf32_sustained"] * world)

    # ---- kernel launch census of one step (CUPTI via torch.profiler; not timed): kernels inside the replayed graph count
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
            mine = [e for e in ev if any(s in e.name for s in ("conv_gemm_kernel", "conv_gemm2_kernel", "conv_wgrad_kernel", "kernel<", "_kernel"))
                    and "at::native" not in e.name and "nccl" not in e.name.lower()]
            line["gpu_launches"] = len(mine)
            line["gpu_launches_all"] = len(ev)
            tot = sum(e.device_time for e in ev) or 1.0
            agg = {}
            for e in ev:
                a = agg.setdefault(_kernel_key(e.name), [0, 0.0])
                a[0] += 1
                a[1] += e.device_time
            gemm = sum(t for k, (n, t) in agg.items() if "conv_gemm" in k or "conv_wgrad" in k)
            line["gemm_share_of_step"] = gemm / tot
            line["top_kernels_by_time"] = [dict(kernel=k, launches=n, ms=t / 1e3, share=t / tot)
                                           for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]]
            if args.breakdown:
                agg2 = {}
                for e in ev:
                    a = agg2.setdefault(e.name[:110], [0, 0.0])
                    a[0] += 1
                    a[1] += e.device_time
                with open(args.breakdown, "w") as f:
                    f.write("# one training step, batch-8 960x1280, torch.profiler (CUPTI) device times\n")
                    f.write("%-112s %6s %10s %6s\n" % ("kernel", "calls", "total_us", "share"))
                    for k, (n, t) in sorted(agg2.items(), key=lambda kv: -kv[1][1]):
                        f.write("%-112s %6d %10.1f %5.1f%%\n" % (k, n, t, 100.0 * t / tot))
        except Exception as ex:      # noqa: BLE001
            line["gpu_launches"] = None
            line["profiler_error"] = str(ex)[:200]

    # the eager (no graph) launch path of the same step, for reference
    if st["graphed"] is not None and not args.no_extras:
        from tinyfaces_b200.trainer import train_step_flat
        eager = timed_steps(lambda: train_step_flat(model, crit, opt, st["dev"][0], st["dev"][1].clone(), st["dev"][2], group), 5, 2, world, dev)
        line["eager_ms_per_step"] = eager / 5
    del st, model, crit, opt, step_resident
    torch.cuda.empty_cache()

    extras = not args.no_extras
    # ---- BASELINE configs[3] / configs[4] (every rank takes part)
    if extras:
        _PROGRESS["phase"] = "cfg3"
        try:
            rep = cfg3_report(dev, world, rank, group, max(5, min(args.steps, 20)))
            if rank == 0:
                line["cfg3"] = rep
        except Exception as ex:  # noqa: BLE001
            line["cfg3"] = dict(error=str(ex)[:300])
        if not args.no_inference:
            _PROGRESS["phase"] = "cfg4"
            try:
                rep = cfg4_report(dev, world, rank, group, args.nms_n)
                if rank == 0:
                    line["cfg4"] = rep
            except Exception as ex:  # noqa: BLE001
                line["cfg4"] = dict(error=str(ex)[:300])

    _PROGRESS["phase"] = "single-GPU extras (parity mode, rooflines, NMS, inference, CPU baseline)"
    if rank == 0 and extras:
        # ---- the arithmetic that meets north_star's 1e-3: 3xTF32 (`parity`), timed at the same shape, and both modes' error
        #      against the reference's golden output at this shape
        if world == 1:
            try:
                fe = golden_forward_error(dev, "fast")
                pe = golden_forward_error(dev, "parity")
                ps = build_step(dev, B_PER_GPU, H_IMG, W_IMG, "parity", rank, None)
                pms = timed_steps(ps["step"], 5, 2, 1, dev) / 5
                line["parity_mode"] = dict(arithmetic="3xTF32 (hi*hi + lo*hi + hi*lo), fp32 accumulate / storage", ms_per_step=pms,
                                           images_per_s=B_PER_GPU / (pms / 1e3), cuda_graph=ps["graphed"] is not None,
                                           out_rel_err=pe, tolerance=1e-3, within_tolerance=pe["max_rel"] < 1e-3 and pe["l2_rel"] < 1e-3,
                                           error_reference="tests/golden/cfg2_b8_fwd.npz: the reference's own 8x3x960x1280 train-mode forward")
                line["fast_out_rel_err"] = fe
                del ps
                torch.cuda.empty_cache()
                # the same forward with the TF32 backward of `fast` (precision="mixed"): outputs identical to `parity`
                ms_ = build_step(dev, B_PER_GPU, H_IMG, W_IMG, "mixed", rank, None)
                mms = timed_steps(ms_["step"], 5, 2, 1, dev) / 5
                line["mixed_mode"] = dict(arithmetic="forward 3xTF32 (= parity: same score map), backward GEMMs 1xTF32 on the hi parts",
                                          ms_per_step=mms, images_per_s=B_PER_GPU / (mms / 1e3), cuda_graph=ms_["graphed"] is not None,
                                          gradient_gate="tests/test_gpu_baseline_shapes.py, tests/test_gpu_model.py: weight gradients "
                                                        "within rel-L2 3e-2 of the reference's fp32 autograd, like parity")
                del ms_
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001
                line["parity_mode"] = dict(error=str(ex)[:300])
        # ---- rooflines: every layer-3 GEMM class alone; the headline `roofline` is the class with the largest time per step
        try:
            rows = roofline_classes(dev, pk)
            line["roofline_classes"] = rows
            ew = elementwise_classes(dev, pk)
            line["roofline_elementwise"] = ew
            traffic = {}
            try:
                with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                    traffic = json.load(f)
            except Exception:  # noqa: BLE001
                pass
            # ---- the dominant kernel of the step (top_kernels_by_time): since the 3x3 weight gradient got its second accumulator
            #      tile, the largest share belongs to the BatchNorm-backward pair bn_bwd_apply_kernel + colreduce_kernel<1> (~31 % of the
            #      kernel time), an HBM-bound pass pair.  Its figure: the launch-weighted aggregate over the layer-3 shapes, each
            #      timed alone with CUDA events; peak = the measured HBM copy bandwidth (burst: the kernels are timed alone).
            bw = [r for r in ew if r["name"].startswith("BN backward")]
            nbytes = sum(r["algorithmic_mb"] * 1e6 * r["launches_per_step"] for r in bw)
            secs = sum(r["us"] * 1e-6 * r["launches_per_step"] for r in bw)
            launches = sum(r["launches_per_step"] for r in bw)
            line["roofline"] = dict(bound="hbm", kernel="bn_bwd_apply_kernel + colreduce_kernel<1> (BatchNorm backward, layer-3 shapes: M=%d pixels, C=1024 / 256)"
                                                        % (B_PER_GPU * (H_IMG // 16) * (W_IMG // 16)),
                                    achieved=nbytes / secs / 1e9, peak=pk["hbm_gbs"], unit="GB/s", frac=nbytes / secs / 1e9 / pk["hbm_gbs"],
                                    traffic=traffic.get("bn_backward_dram_bytes_per_launch"),
                                    traffic_source="profiles/roofline_traffic.json (ncu --set full: dram__bytes_read + dram__bytes_write of the two kernels, launch-weighted mean)",
                                    algorithmic_bytes_per_launch=nbytes / launches, us_per_launch=secs / launches * 1e6,
                                    launches_per_step=launches, peak_source=pk["source"],
                                    why="largest share of the step's kernel time (top_kernels_by_time); 20.25 algorithmic bytes per element "
                                        "(DESIGN.md section 4); the tensor-core kernels are in roofline_tensor / roofline_classes")
            # ---- the tensor-bound kernel with the largest share: conv_wgrad_kernel<256> (launch-weighted over its layer-3 classes)
            wg = [r for r in rows if r["kernel"].startswith("conv_wgrad")]
            flops = sum(r["algorithmic_gflop"] * 1e9 * r["launches_per_step"] for r in wg)
            secs = sum(r["us_per_launch"] * 1e-6 * r["launches_per_step"] for r in wg)
            launches = sum(r["launches_per_step"] for r in wg)
            line["roofline_tensor"] = dict(bound="tensor", kernel="conv_wgrad_kernel<256, 1|2> (layer-3 weight gradients: 3x3 256->256, 1x1 256<->1024)",
                                           achieved=flops / secs / 1e12, peak=pk["tf32_burst"], unit="TFLOP/s", frac=flops / secs / 1e12 / pk["tf32_burst"],
                                           traffic=traffic.get("conv_wgrad_dram_bytes_per_launch"), algorithmic_flops_per_launch=flops / launches,
                                           us_per_launch=secs / launches * 1e6, launches_per_step=launches,
                                           note="the 3x3 class runs at 0.9 of the TF32 burst peak (ncu: 78.7 % tensor-pipe active); the 1x1 classes "
                                                "stream 196 MB per 20 GFLOP and sit at ~73 % of the HBM peak instead (ncu: 56 % tensor pipe)")
        except Exception as ex:  # noqa: BLE001
            line["roofline"] = dict(error=str(ex)[:300])
        # ---- NMS boxes/s (the second half of BASELINE.json's metric)
        try:
            line["nms"] = nms_report(dev, args.nms_n, pk, with_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as ex:  # noqa: BLE001
            line["nms"] = dict(error=str(ex)[:300])
        try:
            line["nms_by_n"] = nms_by_n(dev)
        except Exception as ex:  # noqa: BLE001
            line["nms_by_n"] = dict(error=str(ex)[:300])
        # ---- training-target generation: 63x63x25 cells x 12 ground-truth boxes
        try:
            from tinyfaces_b200 import inference_bench as ib
            from tinyfaces_b200.targets import DataProcessor
            tp = DataProcessor((500, 500), (63, 63), 0.7, 0.3, ib.load_templates()[:, :4], rf=ib.RF, device=dev, jitter="device")
            r = np.random.RandomState(0)
            xy = r.rand(12, 2) * 400
            gt = np.concatenate([xy, xy + 10 + r.rand(12, 2) * 120], axis=1)
            pad = tp.get_padding([10, 10, 490, 490])
            t_t = _event_time(lambda: tp.get_heatmaps_device(gt, pad), 10)
            line["targets"] = dict(workload="DataProcessor.get_heatmaps: 63x63x25 cells x 12 boxes (incl. the host->device copies of boxes / mask)",
                                   gpu_ms=t_t * 1e3, reference_note="the reference's compute_dense_overlap is a 4-deep Python loop: ~0.15 s per box here")
        except Exception as ex:  # noqa: BLE001
            line["targets"] = dict(error=str(ex)[:200])
        # ---- pyramid inference (BASELINE.json configs[2])
        if not args.no_inference and world == 1:
            try:
                from tinyfaces_b200 import inference_bench
                imodel = inference_bench.make_calibrated_model(dev)      # fresh weights with calibrated BN statistics
                line["inference"] = inference_bench.run(imodel, base=1250, target_candidates=args.nms_n)
                del imodel
            except Exception as ex:  # noqa: BLE001
                line["inference"] = dict(error=str(ex)[:300])
        # ---- CPU baseline beside it (bounded sample)
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, cores, sample = cpu_reference_step_rate(steps=2, warmup=1)
                line["cpu_baseline"] = dict(value=v, unit="images/s", cores=cores, kind="port", sample=sample)
                if isinstance(line.get("nms"), dict) and "cpu" in line["nms"]:
                    line["cpu_baseline"]["nms_boxes_per_s"] = line["nms"]["cpu"].get("boxes_per_s")
                    line["cpu_baseline"]["nms_n"] = args.nms_n
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = dict(error=str(ex)[:200])
    if rank == 0:
        print(json.dumps(line), flush=True)
    _PROGRESS["line"] = None                      # the line is out: the watchdog must not print a second one
    sys.stdout.flush()
    if world > 1:
        # leave together, then exit without tearing the process group down: destroy_process_group() can block for minutes
        # while CUDA graphs that captured NCCL kernels are still being released
        try:
            dist.barrier()
        except Exception:  # noqa: BLE001
            pass
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
