"""profiles/roofline_traffic.json from the `ncu --set full` summaries under profiles/ (prof_*_<tag>.txt, written by
tools/summarize_ncu.py): DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the kernels bench.py reports
rooflines for, launch-weighted over the layer-3 shape classes.   python tools/make_roofline_traffic.py r2"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def parse(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return []
    out, cur = [], None
    for line in open(p):
        if line.startswith("== "):
            cur = dict(kernel=line[3:].split("(")[0].strip(), grid=re.search(r"grid \((\d+)", line).group(1))
            out.append(cur)
        elif cur is not None:
            m = re.match(r"(\S+)\s+([0-9.,]+)\s+(\S+)", line)
            if m:
                v = float(m.group(2).replace(",", ""))
                k = m.group(1)
                if k.startswith("dram__bytes"):
                    v *= UNIT.get(m.group(3), 1.0)
                elif k == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(m.group(3), 1.0)
                cur[k] = v
    return out


def dram(r):
    return r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)


res = {"source": "ncu --set full --clock-control none, python bench.py --profile-mode --steps 1 --warmup 1 (in situ, one eager training step, "
                 "batch-8 960x1280); caches are flushed between ncu replays, so these are cold-L2 figures"}
wg = parse("prof_wgrad_%s.txt" % tag)
c33 = [r for r in wg if "conv_wgrad_kernel<256, 2>" in r["kernel"]]
c11 = [r for r in wg if "conv_wgrad_kernel<256, 1>" in r["kernel"]]
if c33 and c11:
    m33, m11 = sum(map(dram, c33)) / len(c33), sum(map(dram, c11)) / len(c11)
    res["wgrad 3x3 256->256"] = dict(dram_bytes_per_launch=m33, algorithmic_bytes=2 * 38400 * 256 * 4 + 256 * 2304 * 4,
                                     us=sum(r["gpu__time_duration.sum"] for r in c33) / len(c33),
                                     tensor_pipe_active_pct=sum(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for r in c33) / len(c33))
    res["wgrad 1x1 256<->1024"] = dict(dram_bytes_per_launch=m11, algorithmic_bytes=38400 * (256 + 1024) * 4 + 256 * 1024 * 4,
                                       us=sum(r["gpu__time_duration.sum"] for r in c11) / len(c11),
                                       tensor_pipe_active_pct=sum(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for r in c11) / len(c11))
    res["conv_wgrad_dram_bytes_per_launch"] = (22 * m33 + 90 * m11) / 112.0
ew = parse("prof_ewbwd_%s.txt" % tag)
ap = {}
for r in ew:
    key = ("apply" if "bn_bwd_apply" in r["kernel"] else "reduce", "big" if dram(r) > 250e6 else "small")
    ap.setdefault(key, []).append(r)
if all(k in ap for k in (("apply", "big"), ("reduce", "big"), ("apply", "small"), ("reduce", "small"))):
    mean = lambda rs: sum(map(dram, rs)) / len(rs)                       # noqa: E731
    big = mean(ap[("apply", "big")]) + mean(ap[("reduce", "big")])
    small = mean(ap[("apply", "small")]) + mean(ap[("reduce", "small")])
    res["BN backward C=1024"] = dict(dram_bytes_per_launch_pair=big, algorithmic_bytes=20.25 * 38400 * 1024)
    res["BN backward C=256"] = dict(dram_bytes_per_launch_pair=small, algorithmic_bytes=20.25 * 38400 * 256)
    res["bn_backward_dram_bytes_per_launch"] = (23 * big + 46 * small) / 69.0
gm = parse("prof_gemm_%s.txt" % tag)
if gm:
    res["conv_gemm captures"] = [dict(kernel=r["kernel"][:60], grid=r["grid"], us=r.get("gpu__time_duration.sum"), dram_bytes=dram(r),
                                      tensor_pipe_active_pct=r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")) for r in gm]
with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps({k: v for k, v in res.items() if not isinstance(v, list)}, indent=1)[:1500])
