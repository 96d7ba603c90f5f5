#!/bin/bash
# GPU call 3: smem-free reductions, statistics-epilogue timing, in-situ ncu capture of the layer-3 GEMMs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c3_pytest_model.log 2>&1
tail -3 gpurun_out/c3_pytest_model.log
timeout 400 python tools/gemm_probe.py > gpurun_out/c3_probe.log 2>&1
cut -c1-200 gpurun_out/c3_probe.log
rm -f gpurun_out/ab_step.jsonl
timeout 900 python tools/ab_step.py "mode0=5:1" "mode1=5:2" "default=" "default_noprio=6:1" > gpurun_out/c3_ab.log 2>&1
cut -c1-330 gpurun_out/c3_ab.log
timeout 300 python tools/timeline.py timeline_c3.csv > gpurun_out/c3_timeline.log 2>&1
tail -1 gpurun_out/c3_timeline.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel --launch-skip 40 --launch-count 3 \
    -f -o gpurun_out/prof_c3 python bench.py --profile-mode --steps 1 --warmup 0 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_ncu.log
ls -la gpurun_out/*.ncu-rep
