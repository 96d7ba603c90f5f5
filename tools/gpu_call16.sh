#!/bin/bash
# GPU call 16: TMA residual epilogue -- inference conv3+BN+shortcut+ReLU in one kernel; training dgrad1 variant (flag 8)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c16_pytest_model.log 2>&1
tail -3 gpurun_out/c16_pytest_model.log
timeout 600 python tools/bench_inference.py > gpurun_out/c16_inference.log 2>&1
tail -1 gpurun_out/c16_inference.log | cut -c1-900
timeout 600 python tools/bench_inference.py --flags 12:1 > gpurun_out/c16_inference_old.log 2>&1
tail -1 gpurun_out/c16_inference_old.log | cut -c1-900
rm -f gpurun_out/ab_step.jsonl
timeout 600 python tools/ab_step.py "default=" "res_tma=8:1" > gpurun_out/c16_ab.log 2>&1
cut -c1-330 gpurun_out/c16_ab.log
