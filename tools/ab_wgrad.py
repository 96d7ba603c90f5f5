"""A/B of the weight-gradient kernel: one accumulator tile per CTA (tf_debug_set(10, 1)) vs two (default), isolated per
layer-3 class and in the whole step (graph replay)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
from tinyfaces_b200 import ops
from tinyfaces_b200._lib import lib
import bench
dev = torch.device("cuda:0")
pk = bench.peaks()
def classes(tag):
    for name, cin, cout, k in (("3x3 256->256", 256, 256, 3), ("1x1 256->1024", 256, 1024, 1), ("1x1 1024->256", 1024, 256, 1), ("1x1 128->512 (layer2)", 128, 512, 1)):
        B, H, W = (8, 60, 80) if "layer2" not in name else (8, 120, 160)
        x = torch.randn(B, H, W, cin, device=dev); dy = torch.randn(B, H, W, cout, device=dev)
        dw = torch.zeros(cout, k * k, cin, device=dev)
        t = bench._event_time(lambda: ops.conv2d_wgrad_nhwc(x, dy, k, out=dw), 40)
        fl = 2.0 * B * H * W * cin * cout * k * k
        # correctness of the new tiling against torch (TF32-exact operands)
        xt = (x.view(torch.int32) & ~0x1FFF).view(torch.float32); dyt = (dy.view(torch.int32) & ~0x1FFF).view(torch.float32)
        dw.zero_(); ops.conv2d_wgrad_nhwc(xt, dyt, k, out=dw)
        if k == 1:
            ref = dyt.reshape(-1, cout).double().t() @ xt.reshape(-1, cin).double()
            err = float((dw.reshape(cout, cin).double() - ref).abs().max() / ref.abs().max())
        else:
            err = None
        print(json.dumps(dict(tag=tag, cls=name, us=t * 1e6, tflops=fl / t / 1e12, frac=fl / t / 1e12 / pk["tf32_burst"], err=err)), flush=True)
lib().tf_debug_set(10, 1); classes("MT=1")
lib().tf_debug_set(10, 0); classes("MT=2")
for flag, tag in ((1, "MT=1"), (0, "MT=2")):
    lib().tf_debug_set(10, flag)
    st = bench.build_step(dev, 8, 960, 1280, "fast", 0, None)
    ms = bench.timed_steps(st["step"], 10, 3, 1, dev) / 10
    print(json.dumps(dict(tag=tag, step_ms=ms, graph=st["graphed"] is not None)), flush=True)
    del st; torch.cuda.empty_cache()
