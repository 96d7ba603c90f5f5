#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "tail_split_k or stride2_dgrad or (conv2d_nhwc_vs_torch and 1024) or (wgrad_vs_torch and 256)" > gpurun_out/sanitizer_ops.log 2>&1
echo "ops rc=$?"; tail -6 gpurun_out/sanitizer_ops.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "layerwise or eval_forward" > gpurun_out/sanitizer_model.log 2>&1
echo "model rc=$?"; tail -6 gpurun_out/sanitizer_model.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "train_loop or pipelined" > gpurun_out/sanitizer_train.log 2>&1
echo "train rc=$?"; tail -6 gpurun_out/sanitizer_train.log
grep -c "ERROR SUMMARY" gpurun_out/sanitizer_*.log
