#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/gemm_probe.py > gpurun_out/c10_probe.log 2>&1
cut -c1-220 gpurun_out/c10_probe.log
rm -f gpurun_out/ab_step.jsonl
timeout 400 python tools/ab_step.py "default=" "pf=10:1" > gpurun_out/c10_ab.log 2>&1
cut -c1-200 gpurun_out/c10_ab.log
