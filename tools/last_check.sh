#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/last_pytest.log 2>&1
tail -2 gpurun_out/last_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_last.json'))
print({k:d[k] for k in ('value','ms_per_step','loss','gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])
print('roofline', d['roofline']['frac'], 'nms', d['nms']['boxes_per_s'], 'inference ms', d['inference']['total_gpu_ms'], 'cpu', d['cpu_baseline']['value'])
PY
