#!/bin/bash
# GPU call 6: rewritten im2col / max-pool / head kernels; full GPU test suite; step timing + timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1
tail -4 gpurun_out/c6_pytest.log
rm -f gpurun_out/ab_step.jsonl
timeout 400 python tools/ab_step.py "default=" > gpurun_out/c6_ab.log 2>&1
cut -c1-330 gpurun_out/c6_ab.log
timeout 300 python tools/timeline.py timeline_c6.csv > gpurun_out/c6_timeline.log 2>&1
tail -1 gpurun_out/c6_timeline.log
