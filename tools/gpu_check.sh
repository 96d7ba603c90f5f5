#!/bin/bash
# quick GPU check: the whole -m gpu suite + smoke (TAG names the logs)
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_$TAG.log 2>&1
tail -25 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
