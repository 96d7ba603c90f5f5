#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (NMS sweep / grid / resolve, BN / pool / head entry points, two-tile wgrad,
# target generation, SGD, flat backward).  Logs: gpurun_out/sanitizer_r2_*.log, gpurun_out/racecheck_r2_*.log
mkdir -p gpurun_out
MC="compute-sanitizer --tool memcheck --error-exitcode 1"
run() { # name, timeout, tool..., -- pytest args
  local name=$1 t=$2; shift 2
  timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; grep -h "passed\|failed\|ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/$name.log | tail -2
}
run sanitizer_r2_nms 300 $MC python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "nms and not large_n"
run sanitizer_r2_kernels 240 $MC python -m pytest tests/test_gpu_kernels.py tests/test_gpu_targets.py -m gpu -x -q
run sanitizer_r2_wgrad 200 $MC python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "wgrad_vs_torch or 3xtf32"
run sanitizer_r2_training 300 $MC python -m pytest tests/test_gpu_training.py -m gpu -x -q -k "sgd_kernel or flat_sgd or flat_store or autograd_free"
RC="compute-sanitizer --tool racecheck --racecheck-report analysis"
run racecheck_r2_nms 240 $RC python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "nms_golden or nms_known or nms_negative"
run racecheck_r2_kernels 240 $RC python -m pytest tests/test_gpu_kernels.py tests/test_gpu_targets.py -m gpu -x -q -k "bn_train or maxpool or head_upsample or golden"
