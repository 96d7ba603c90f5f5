#!/bin/bash
# tests + smoke + bench (TAG names the outputs); optional second arg "ncu" adds the launch list and the backward-elementwise capture
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 800 python bench.py --breakdown gpurun_out/step_breakdown_$TAG.txt > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
if [ "$2" == "ncu" ]; then
  NCU="ncu --clock-control none"
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_r2.csv \
      python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_launches_r2.log 2>&1; tail -1 gpurun_out/ncu_launches_r2.log
  timeout 400 $NCU --set full --import-source on -k "regex:bn_bwd_apply_kernel|colreduce_kernel" --launch-skip 238 --launch-count 8 -f -o gpurun_out/prof_ewbwd_r2 \
      python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_ewbwd_r2.log 2>&1; tail -1 gpurun_out/ncu_ewbwd_r2.log
fi
