#!/bin/bash
# GPU call 4: compact epilogue (explicit LDS/STS, FUSED/SPATIAL instantiations), wgrad smem overlay, EW occupancy cap A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv2d" > gpurun_out/c4_pytest_ops.log 2>&1
tail -3 gpurun_out/c4_pytest_ops.log
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/c4_pytest_model.log 2>&1
tail -3 gpurun_out/c4_pytest_model.log
timeout 400 python tools/gemm_probe.py > gpurun_out/c4_probe.log 2>&1
cut -c1-200 gpurun_out/c4_probe.log
rm -f gpurun_out/ab_step.jsonl
timeout 900 python tools/ab_step.py "default=" "ew7=env:TF_EW_BLOCKS_PER_SM:7" "ew7_noprio=env:TF_EW_BLOCKS_PER_SM:7,6:1" "ew7_mode1=env:TF_EW_BLOCKS_PER_SM:7,5:2" "ew7_mode0=env:TF_EW_BLOCKS_PER_SM:7,5:1" "ew4=env:TF_EW_BLOCKS_PER_SM:4" > gpurun_out/c4_ab.log 2>&1
cut -c1-330 gpurun_out/c4_ab.log
TF_EW_BLOCKS_PER_SM=7 timeout 300 python tools/timeline.py timeline_c4.csv > gpurun_out/c4_timeline.log 2>&1
tail -1 gpurun_out/c4_timeline.log
