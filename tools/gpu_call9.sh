#!/bin/bash
# GPU call 9: weight packing off the critical path, finalize, foreach counters; NMS check; timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1
tail -3 gpurun_out/c9_pytest.log
rm -f gpurun_out/ab_step.jsonl
timeout 400 python tools/ab_step.py "default=" > gpurun_out/c9_ab.log 2>&1
cut -c1-330 gpurun_out/c9_ab.log
timeout 300 python tools/timeline.py timeline_c9.csv > gpurun_out/c9_timeline.log 2>&1
tail -1 gpurun_out/c9_timeline.log
timeout 300 python tools/profile_nms.py 100000 1000000 > gpurun_out/c9_nms.log 2>&1
cat gpurun_out/c9_nms.log
