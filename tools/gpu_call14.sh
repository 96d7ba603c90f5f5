#!/bin/bash
# GPU call 14: one statistics row per CTA, phase-decomposed stride-2 dgrad
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv2d" > gpurun_out/c14_pytest_ops.log 2>&1
tail -4 gpurun_out/c14_pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c14_pytest_model.log 2>&1
tail -4 gpurun_out/c14_pytest_model.log
rm -f gpurun_out/ab_step.jsonl
timeout 600 python tools/ab_step.py "default=" "old_s2=11:1" > gpurun_out/c14_ab.log 2>&1
cut -c1-330 gpurun_out/c14_ab.log
