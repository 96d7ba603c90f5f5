#!/bin/bash
# precision="mixed": gradient parity at the test shapes and the step time next to parity / fast
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_model.py tests/test_gpu_baseline_shapes.py -m gpu -q -k "mixed or (parity and (cfg1_train or cfg2_b1 or large_batch or reference_golden))" > gpurun_out/pytest_mixed.log 2>&1; tail -5 gpurun_out/pytest_mixed.log
timeout 300 python - <<'PY' 2>&1 | tail -8
import sys, json, torch
sys.path.insert(0, "tiny-faces-pytorch_b200"); sys.path.insert(0, ".")
import bench
dev = torch.device("cuda:0")
for prec in ("mixed", "parity", "fast", "mixed"):
    st = bench.build_step(dev, 8, 960, 1280, prec, 0, None)
    ms = bench.timed_steps(st["step"], 6, 2, 1, dev) / 6
    print(json.dumps(dict(precision=prec, graph=st["graphed"] is not None, err=st["graph_error"], step_ms=round(ms, 2), mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 1))), flush=True)
    del st; torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
PY
grep -h "mixed\|parity" gpurun_out/model_parity.jsonl gpurun_out/baseline_shape_parity.jsonl | tail -14 | cut -c1-900
