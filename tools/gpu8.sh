#!/bin/bash
# N-GPU bench under torchrun with an outer timeout; prints the interesting keys (usage: gpu8.sh <N>)
N=${1:-8}
mkdir -p gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 10 --warmup 3 --deadline 520 > gpurun_out/bench_${N}gpu_r2.json 2> gpurun_out/bench_${N}gpu_r2.err
echo "bench rc=$?"; grep -v "^W\|^\[W\|Warning\|warn\|^\*\|OMP_NUM" gpurun_out/bench_${N}gpu_r2.err | tail -5
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_${N}gpu_r2.json") if l.startswith("{")][-1]
    for k in ("value","ms_per_step","eager_ms_per_step","watchdog","cfg3","cfg4"):
        print(k, json.dumps(d.get(k))[:1200])
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "graph", d["config"].get("cuda_graph"), d["config"].get("graph_error"))
except Exception as ex: print("no bench json:", ex)
PY
