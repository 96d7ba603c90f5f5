"""profiles/sass_summary_<tag>.txt: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use
(cuobjdump -sass of the built library; no GPU needed).   python tools/sass_summary.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tiny-faces-pytorch_b200", "tinyfaces_b200", "libtinyfaces_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
MN = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "DADD", "DMUL", "DFMA"]
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["instr"] += 1
        for k in MN:
            if op == k or op.startswith(k + "."):
                if k == "UTCHMMA" and ".2CTA" in op:
                    continue
                counts[kern][k] += 1
                break
with open(os.path.join(ROOT, "profiles", "sass_summary_%s.txt" % tag), "w") as f:
    f.write("# cuobjdump -sass libtinyfaces_b200.so (sm_100a), per kernel: SASS instruction count and the tcgen05 / TMA mnemonics\n")
    f.write("# UTCHMMA = tcgen05.mma (.2CTA: cta_group::2), UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, LDTM = tcgen05.ld,\n")
    f.write("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0), D* = fp64 (NMS / decode / targets)\n")
    f.write("%-78s %6s " % ("kernel", "instr") + " ".join("%12s" % k for k in MN) + "\n")
    for k, c in counts.items():
        f.write("%-78s %6d " % (k[:78], c["instr"]) + " ".join("%12d" % c[m] for m in MN) + "\n")
        for m in MN:
            total[m] += c[m]
    f.write("%-78s %6d " % ("TOTAL (%d kernels)" % len(counts), sum(c["instr"] for c in counts.values())) + " ".join("%12d" % total[m] for m in MN) + "\n")
print("wrote profiles/sass_summary_%s.txt: %d kernels, UTCHMMA %d (+%d 2CTA), HMMA %d" % (tag, len(counts), total["UTCHMMA"], total["UTCHMMA.2CTA"], total["HMMA"]))
