#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-inference --no-cpu-baseline > gpurun_out/bench_2gpu_c13.json 2> gpurun_out/bench_2gpu_c13.err
cat gpurun_out/bench_2gpu_c13.json | cut -c1-1500; tail -3 gpurun_out/bench_2gpu_c13.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multi_gpu.py > gpurun_out/check_2gpu_c13.log 2>&1
tail -6 gpurun_out/check_2gpu_c13.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_2gpu_c13.json 2>&1
tail -2 gpurun_out/bench_ref_2gpu_c13.json | cut -c1-300
