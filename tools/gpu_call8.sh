#!/bin/bash
# GPU call 8: L2-aware iteration order of the BN passes (A/B)
mkdir -p gpurun_out
rm -f gpurun_out/ab_step.jsonl
timeout 900 python tools/ab_step.py "default=" "apply_desc=9:1" "colred_desc=9:2" "both_desc=9:3" > gpurun_out/c8_ab.log 2>&1
cut -c1-330 gpurun_out/c8_ab.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q > gpurun_out/c8_pytest.log 2>&1
tail -3 gpurun_out/c8_pytest.log
