#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1
tail -4 gpurun_out/c12_pytest.log
timeout 1200 python bench.py > gpurun_out/bench_c12.json 2> gpurun_out/bench_c12.err
cat gpurun_out/bench_c12.json; tail -3 gpurun_out/bench_c12.err
