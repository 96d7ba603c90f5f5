#!/bin/bash
# 2-GPU validation, every stage under its own timeout, logs in gpurun_out/
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for st in inference dp graph; do
  timeout -k 5 ${2:-150} $TR --master-port $((29600 + RANDOM % 300)) tools/check_multi_gpu.py $st > gpurun_out/mg_$st.log 2>&1
  echo "stage $st rc=$?"; grep -v "^W\|^\[W\|Warning\|warn" gpurun_out/mg_$st.log | tail -${1:-6}
done
timeout -k 5 420 $TR --master-port $((29900 + RANDOM % 90)) bench.py --gpus 2 --steps 10 --warmup 3 --deadline 380 > gpurun_out/bench_2gpu_r2a.json 2> gpurun_out/bench_2gpu_r2a.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_2gpu_r2a.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_2gpu_r2a.json") if l.startswith("{")][-1]
    for k in ("value","ms_per_step","eager_ms_per_step","watchdog","cfg3","cfg4"):
        print(k, json.dumps(d.get(k))[:900])
    print("config", json.dumps(d["config"])[:500]); print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as ex: print("no bench json:", ex)
PY
