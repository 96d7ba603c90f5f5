"""Per-kernel time of tf_nms at a few N (torch.profiler), to find the bottleneck stage."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-faces-pytorch_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
from tinyfaces_b200 import ops, synthetic

for n in [int(a) for a in sys.argv[1:]] or [100000]:
    b, s = synthetic.boxes(n, seed=0)
    b, s = b.cuda(), s.cuda()
    for _ in range(2):
        ops.nms_device(b, s, 0.3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keep, cnt = ops.nms_device(b, s, 0.3)
    e1.record()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ops.nms_device(b, s, 0.3)
        torch.cuda.synchronize()
    agg = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            a = agg.setdefault(e.name[:70], [0, 0.0])
            a[0] += 1
            a[1] += e.device_time
    print("N=%d kept=%d total %.2f ms -> %.2f M boxes/s" % (n, int(cnt.item()), e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e3))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
        print("   %-72s %4d %10.1f us" % (k, c, t))
