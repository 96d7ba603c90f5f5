#!/bin/bash
# usage: bash tools/quick_bench_ngpu.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','loss')}, d['e2e']['value'])
PY
tail -2 gpurun_out/bench_${N}gpu.err
