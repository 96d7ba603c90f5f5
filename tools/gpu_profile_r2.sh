#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of one eager step + full captures of the weight-gradient, elementwise and GEMM kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_launches_r2.log 2>&1; tail -1 gpurun_out/ncu_launches_r2.log
timeout 400 $NCU --set full --import-source on -k regex:conv_wgrad_kernel --launch-skip 130 --launch-count 6 -f -o gpurun_out/prof_wgrad_r2 \
    python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_wgrad_r2.log 2>&1; tail -1 gpurun_out/ncu_wgrad_r2.log
timeout 400 $NCU --set full --import-source on -k "regex:bn_bwd_apply_kernel|colreduce_kernel" --launch-skip 238 --launch-count 8 -f -o gpurun_out/prof_ewbwd_r2 \
    python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_ewbwd_r2.log 2>&1; tail -1 gpurun_out/ncu_ewbwd_r2.log
timeout 400 $NCU --set full --import-source on -k "regex:bn_bwd_apply_kernel|colreduce_kernel|bn_apply_kernel" --launch-skip 330 --launch-count 9 -f -o gpurun_out/prof_ew_r2 \
    python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_ew_r2.log 2>&1; tail -1 gpurun_out/ncu_ew_r2.log
timeout 400 $NCU --set full --import-source on -k regex:conv_gemm --launch-skip 330 --launch-count 9 -f -o gpurun_out/prof_gemm_r2 \
    python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_gemm_r2.log 2>&1; tail -1 gpurun_out/ncu_gemm_r2.log
ls -la gpurun_out/*.ncu-rep
