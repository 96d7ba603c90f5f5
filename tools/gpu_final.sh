#!/bin/bash
# Round measurement: full GPU tests, both bench arms, ncu launch list + full capture of the layer-3 GEMMs, breakdown.
TAG=${1:-r1b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_$TAG.log 2>&1
tail -3 gpurun_out/final_pytest_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke_$TAG.log 2>&1
tail -1 gpurun_out/final_smoke_$TAG.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
cat gpurun_out/bench_ref_$TAG.json | cut -c1-400
timeout 1200 python bench.py --breakdown gpurun_out/step_breakdown_$TAG.txt > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip 40 --launch-count 3 \
    -f -o gpurun_out/prof_gemm_$TAG python bench.py --profile-mode --steps 1 --warmup 0 > gpurun_out/ncu_gemm_$TAG.log 2>&1
tail -2 gpurun_out/ncu_gemm_$TAG.log
ls -la gpurun_out/*.ncu-rep
