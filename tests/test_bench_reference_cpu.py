"""bench.py's reference arm (the oracle port timed on the host cores) runs without a GPU and keeps the driver's JSON
contract: one line, the metric / unit / config of the main arm, `impl`, `cpu_baseline`, an `e2e` that repeats the value."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd, env=None):
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert "fwd+loss+bwd+SGD" in d["metric"] and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "quarter-area" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    lines = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                  "--master-port", "29533", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env=env)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
