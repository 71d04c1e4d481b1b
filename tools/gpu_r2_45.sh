#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "by_class_blocks" > gpurun_out/r2_45_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2_45_pytest.log
tail -4 gpurun_out/r2_45_pytest.log
