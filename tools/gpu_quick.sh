#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "${1:-cross_entropy or seg_ce or wrapper}" > gpurun_out/pytest_quick.log 2>&1; echo "exit $?" >> gpurun_out/pytest_quick.log
timeout 120 python tools/segce_bench.py > gpurun_out/segce.log 2>&1
tail -6 gpurun_out/pytest_quick.log; cat gpurun_out/segce.log
