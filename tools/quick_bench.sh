#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --no-inference --no-cpu-baseline > gpurun_out/bench_c19.json 2> gpurun_out/bench_c19.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c19.json'))
print({k:d[k] for k in ('value','ms_per_step','loss','gpu_launches')}, d['e2e']['value'], d['clocks'])
PY
