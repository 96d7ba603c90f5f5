#!/bin/bash
# GPU call 11: single-sweep NMS (edge list + edge/node relaxation), 2-CTA default for flat K-heavy GEMMs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1
tail -4 gpurun_out/c11_pytest.log
timeout 300 python tools/profile_nms.py 100000 1000000 > gpurun_out/c11_nms.log 2>&1
grep -v Warn gpurun_out/c11_nms.log | cut -c1-120
rm -f gpurun_out/ab_step.jsonl
timeout 400 python tools/ab_step.py "default=" "no2cta=4:2" > gpurun_out/c11_ab.log 2>&1
cut -c1-200 gpurun_out/c11_ab.log
