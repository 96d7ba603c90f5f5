#!/bin/bash
# GPU call 5: max-shared carveout for chain EW kernels, slot-accumulated BN backward reduction; schedule A/B; timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c5_pytest_model.log 2>&1
tail -3 gpurun_out/c5_pytest_model.log
rm -f gpurun_out/ab_step.jsonl
timeout 900 python tools/ab_step.py "default=" "noprio=6:1" "mode1=5:2" "mode0=5:1" "mode0_noprio=5:1,6:1" > gpurun_out/c5_ab.log 2>&1
cut -c1-330 gpurun_out/c5_ab.log
timeout 300 python tools/timeline.py timeline_c5.csv > gpurun_out/c5_timeline.log 2>&1
tail -1 gpurun_out/c5_timeline.log
