"""First-contact probe for the tcgen05 kernels: every configuration runs in its own subprocess with a
timeout (a trapped kernel poisons its CUDA context), results go to gpurun_out/gemm_probe.jsonl."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
sys.path.insert(0, ROOT)

CASES_OLD = [
    ("fprop", 1, 8, 16, 32, 64, 1), ("fprop", 1, 8, 16, 32, 128, 1), ("fprop", 1, 8, 16, 32, 256, 1),
    ("fprop", 2, 17, 23, 64, 128, 1), ("fprop", 1, 30, 40, 256, 256, 1), ("fprop", 1, 8, 16, 32, 64, 3),
    ("fprop", 2, 15, 20, 256, 256, 3),
    ("wgrad", 1, 8, 16, 32, 64, 1), ("wgrad", 1, 8, 16, 128, 128, 1), ("wgrad", 2, 17, 23, 256, 256, 1),
    ("wgrad", 1, 8, 16, 32, 64, 3), ("wgrad", 2, 15, 20, 256, 256, 3),
]


CASES_R1B = [
    ("fprop+2cta", 1, 30, 40, 256, 256, 1), ("fprop+2cta", 2, 15, 20, 1024, 256, 1), ("fprop+2cta", 1, 13, 29, 256, 1024, 1),
    ("fprop+2cta", 2, 15, 20, 256, 256, 3), ("fprop+2cta", 1, 63, 63, 256, 256, 3), ("fprop+2cta", 3, 9, 16, 256, 512, 1),
]
CASES = [
    ("fprop+pf", 2, 15, 20, 1024, 256, 1), ("fprop+pf", 3, 33, 47, 512, 128, 1),
    ("time", 8, 60, 80, 1024, 256, 1), ("time+pf", 8, 60, 80, 1024, 256, 1), ("time+2cta", 8, 60, 80, 1024, 256, 1),
    ("time+2cta+stats", 8, 60, 80, 1024, 256, 1), ("time+2cta+stats", 8, 60, 80, 256, 1024, 1), ("time+stats", 8, 60, 80, 256, 1024, 1),
    ("time", 8, 120, 160, 512, 128, 1), ("time+pf", 8, 120, 160, 512, 128, 1),
    ("time", 8, 30, 40, 1024, 256, 1), ("time+pf", 8, 30, 40, 1024, 256, 1),
]
CASES_OLD2 = [
    ("wgrad", 1, 8, 16, 32, 64, 1), ("wgrad+plain", 1, 8, 16, 32, 64, 1), ("wgrad+plain", 1, 8, 16, 128, 128, 1),
    ("wgrad+plain", 2, 17, 23, 256, 256, 1), ("wgrad+plain", 2, 15, 20, 256, 256, 3),
    ("fprop+acc", 1, 8, 16, 32, 64, 1), ("fprop+acc", 2, 15, 20, 256, 256, 3), ("wgrad", 2, 15, 20, 256, 256, 3),
]


def run_case(kind, B, H, W, Cin, Cout, k):
    import torch
    from tinyfaces_b200 import ops, _lib
    if kind.endswith("+plain"):
        _lib.lib().tf_debug_set(0, 1)
        kind = "wgrad"
    if kind.endswith("+stats"):                   # force the BN-statistics epilogue
        _lib.lib().tf_debug_set(7, 1)
        kind = kind[:-6]
    if kind.endswith("+nosplit"):                 # tail-wave split-K off
        _lib.lib().tf_debug_set(3, 2)
        kind = kind[:-8]
    if kind.endswith("+2cta"):
        _lib.lib().tf_debug_set(4, 1)
        kind = kind[:-5]
    else:
        _lib.lib().tf_debug_set(4, 2)
    if kind == "timew":
        d = torch.device("cuda:0")
        xn = torch.randn(B, H, W, Cin, device=d)
        dyn = torch.randn(B, H, W, Cout, device=d)
        dw = torch.zeros(Cout, k * k, Cin, device=d)
        for _ in range(5):
            ops.conv2d_wgrad_nhwc(xn, dyn, k, out=dw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ops.conv2d_wgrad_nhwc(xn, dyn, k, out=dw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        return dict(kind="timew", shape=[B, H, W, Cin, Cout, k], us=us, tflops=2.0 * B * H * W * Cin * Cout * k * k / us / 1e6,
                    flag=ops.gemm_error_flag())
    if kind == "time":
        d = torch.device("cuda:0")
        xn = torch.randn(B, H, W, Cin, device=d)
        wp = torch.randn(Cout, k * k, Cin, device=d) * 0.02
        y = torch.empty(B, H, W, Cout, device=d)
        for _ in range(5):
            ops.conv2d_nhwc(xn, wp, k, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ops.conv2d_nhwc(xn, wp, k, out=y)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        return dict(kind="time", shape=[B, H, W, Cin, Cout, k], us=us, tflops=2.0 * B * H * W * Cin * Cout * k * k / us / 1e6,
                    flag=ops.gemm_error_flag())
    acc = kind.endswith("+acc")
    if acc:
        _lib.lib().tf_debug_set(1, 1)
        kind = "fprop"
    d = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1)

    def tf32(t):
        i = t.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

    x = tf32(torch.randn(B, Cin, H, W, generator=gen))
    res = dict(kind=kind, shape=[B, H, W, Cin, Cout, k])
    if kind == "fprop":
        w = tf32(torch.randn(Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5)
        ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=k // 2).float()
        y0 = torch.ones((B, H, W, Cout), dtype=torch.float32, device=d) if acc else None
        y = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(d),
                            w.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous().to(d), k, out=y0)
        torch.cuda.synchronize()
        got = y.cpu().permute(0, 3, 1, 2)
        if acc:
            ref = ref + 1
    else:
        dy = tf32(torch.randn(B, Cout, H, W, generator=gen))
        w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
        torch.nn.functional.conv2d(x.double(), w, padding=k // 2).backward(dy.double())
        ref = w.grad.float()
        dw = ops.conv2d_wgrad_nhwc(x.permute(0, 2, 3, 1).contiguous().to(d), dy.permute(0, 2, 3, 1).contiguous().to(d), k)
        torch.cuda.synchronize()
        got = dw.cpu().reshape(Cout, k, k, Cin).permute(0, 3, 1, 2)
    diff = (got - ref).abs()
    res["rel_err"] = diff.max().item() / ref.abs().max().item()
    res["frac_bad"] = (diff > 1e-3 * ref.abs().max()).float().mean().item()
    res["got_absmax"] = got.abs().max().item()
    res["ref_absmax"] = ref.abs().max().item()
    res["flag"] = ops.gemm_error_flag()
    res["nonzero_frac"] = (got != 0).float().mean().item()
    g, r = got.flatten().double(), ref.flatten().double()
    res["corr"] = float((g * r).sum() / ((g.norm() * r.norm()) + 1e-30))
    # a few samples to diagnose layout mistakes
    res["got0"] = got.flatten()[:6].tolist()
    res["ref0"] = ref.flatten()[:6].tolist()
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1:
        args = json.loads(sys.argv[1])
        print("RESULT " + json.dumps(run_case(*args)))
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_probe.jsonl"), "w") as f:
        for c in CASES:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), json.dumps(c)], capture_output=True,
                                   text=True, timeout=180)
                line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                rec = json.loads(line[0][7:]) if line else dict(case=c, rc=r.returncode, err=r.stderr[-1500:])
            except subprocess.TimeoutExpired:
                rec = dict(case=c, timeout=True)
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:400])
