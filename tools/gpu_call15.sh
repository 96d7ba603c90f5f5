#!/bin/bash
# GPU call 15: FUSED epilogue (shift in smem, BN scale folded into the packed weights) -- eval/inference path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv2d" > gpurun_out/c15_pytest_ops.log 2>&1
tail -3 gpurun_out/c15_pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c15_pytest_model.log 2>&1
tail -3 gpurun_out/c15_pytest_model.log
timeout 600 python tools/bench_inference.py > gpurun_out/c15_inference.log 2>&1
tail -5 gpurun_out/c15_inference.log
