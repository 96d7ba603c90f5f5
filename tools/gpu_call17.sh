#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c17_pytest.log 2>&1
tail -3 gpurun_out/c17_pytest.log
timeout 600 python tools/bench_inference.py > gpurun_out/c17_inference.log 2>&1
tail -1 gpurun_out/c17_inference.log | cut -c1-700
rm -f gpurun_out/ab_step.jsonl
timeout 600 python tools/ab_step.py "default=" > gpurun_out/c17_ab.log 2>&1
cut -c1-250 gpurun_out/c17_ab.log
