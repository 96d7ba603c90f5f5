"""Condense ncu outputs from gpurun_out/ into small tracked text files under profiles/.
    python tools/summarize_ncu.py <tag>      (expects gpurun_out/launches_<tag>.csv and/or gpurun_out/prof_*_<tag>.ncu-rep)
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

lp = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(lp):
    lines = open(lp).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for r in csv.DictReader(lines[start:]):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v / 1000 if u.startswith("ns") else (v if u.startswith("us") else v * 1000)
        rows.append((int(r["ID"]), re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""),
                     r["Grid Size"], us))
    idx = [i for i, r in enumerate(rows) if "stem_im2col" in r[1]]
    step = rows[idx[-1]:] if idx else rows
    tot = sum(r[3] for r in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        agg[r[1]][0] += 1
        agg[r[1]][1] += r[3]
    with open(os.path.join(ROOT, "profiles", "launches_%s.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --profile-mode --steps 1 --warmup 1\n")
        f.write("# last training step (batch-8 960x1280): %d launches, %.1f us summed (cold-cache, serialised: compare SHARES)\n" % (len(step), tot))
        f.write("%-60s %6s %10s %6s\n" % ("kernel", "calls", "total_us", "share"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %6d %10.1f %5.1f%%\n" % (k[:60], n, t, 100 * t / tot))
        f.write("\n# 25 longest single launches (id, kernel, grid, us)\n")
        for r in sorted(step, key=lambda r: -r[3])[:25]:
            f.write("%6d %-50s %-14s %9.1f\n" % (r[0], r[1][:50], r[2], r[3]))
    print("wrote profiles/launches_%s.txt" % tag)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_active.avg", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg"]
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*_%s.ncu-rep" % tag))):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, units = rd[0], rd[1]
    ix = {h: i for i, h in enumerate(hdr)}
    name = os.path.basename(rep).replace(".ncu-rep", "")
    with open(os.path.join(ROOT, "profiles", name + ".txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on  (%s)\n" % os.path.basename(rep))
        for r in rd[2:]:
            f.write("\n== %s  grid %s block %s\n" % (r[ix["Kernel Name"]][:100], r[ix["Grid Size"]], r[ix["Block Size"]]))
            for w in WANT:
                if w in ix:
                    f.write("%-72s %16s %s\n" % (w, r[ix[w]], units[ix[w]]))
    print("wrote profiles/%s.txt" % name)
