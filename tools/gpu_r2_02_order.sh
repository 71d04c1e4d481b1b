#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
PCL_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -rA --tb=long -p no:cacheprovider \
   > gpurun_out/r2_02_pytest_full.log 2>&1; echo "exit $?" >> gpurun_out/r2_02_pytest_full.log
timeout 300 python tools/fp32_determinism.py > gpurun_out/r2_02_determinism.log 2>&1; echo "exit $?" >> gpurun_out/r2_02_determinism.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_02_pytest_full.log | tail -20
tail -20 gpurun_out/r2_02_determinism.log
