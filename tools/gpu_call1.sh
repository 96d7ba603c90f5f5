#!/bin/bash
# GPU call 1: 2-CTA kernel probe, wgrad-schedule A/B, model/pipeline parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 python tools/gemm_probe.py > gpurun_out/c1_probe.log 2>&1
timeout 900 python tools/ab_step.py > gpurun_out/c1_ab.log 2>&1
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
tail -3 gpurun_out/c1_pytest.log
cat gpurun_out/c1_ab.log | cut -c1-400
