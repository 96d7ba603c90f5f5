#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c18_pytest.log 2>&1
tail -3 gpurun_out/c18_pytest.log
rm -f gpurun_out/ab_step.jsonl
timeout 400 python tools/ab_step.py "default=" > gpurun_out/c18_ab.log 2>&1
cut -c1-250 gpurun_out/c18_ab.log
