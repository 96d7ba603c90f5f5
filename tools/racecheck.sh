#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "(tail_split_k and 100-100-512) or (conv2d_nhwc_vs_torch and 1024)" > gpurun_out/racecheck_ops.log 2>&1
echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/racecheck_ops.log | head -20
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "eval_forward and fast" > gpurun_out/racecheck_model.log 2>&1
echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/racecheck_model.log | head -20
