#!/bin/bash
# GPU call 2: tail split-K parity + timing, A/B of the step, kernel timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv2d" > gpurun_out/c2_pytest_ops.log 2>&1
tail -5 gpurun_out/c2_pytest_ops.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q > gpurun_out/c2_pytest_model.log 2>&1
tail -5 gpurun_out/c2_pytest_model.log
timeout 600 python tools/gemm_probe.py > gpurun_out/c2_probe.log 2>&1
cut -c1-200 gpurun_out/c2_probe.log
rm -f gpurun_out/ab_step.jsonl
timeout 900 python tools/ab_step.py > gpurun_out/c2_ab.log 2>&1
cut -c1-1200 gpurun_out/c2_ab.log
timeout 300 python tools/timeline.py timeline_c2.csv > gpurun_out/c2_timeline.log 2>&1
tail -2 gpurun_out/c2_timeline.log
