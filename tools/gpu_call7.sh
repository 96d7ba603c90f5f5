#!/bin/bash
# GPU call 7: residual epilogue in the conv1 dgrad (no G tensor), stem scratch behind the regions, fused SGD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1
tail -4 gpurun_out/c7_pytest.log
rm -f gpurun_out/ab_step.jsonl
timeout 600 python tools/ab_step.py "default=" "old_g=8:1" > gpurun_out/c7_ab.log 2>&1
cut -c1-330 gpurun_out/c7_ab.log
timeout 300 python tools/timeline.py timeline_c7.csv > gpurun_out/c7_timeline.log 2>&1
tail -1 gpurun_out/c7_timeline.log
