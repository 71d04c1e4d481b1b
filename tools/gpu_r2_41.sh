#!/bin/bash
# round 2, visit 41 (1 GPU): seg-CE backward with the transposed bilinear weights precomputed once per CTA
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "upsample or seg_ce or wrapper or hook" > gpurun_out/r2_41_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2_41_pytest.log
timeout 300 python tools/segce_bench.py > gpurun_out/r2_41_segce_timing.log 2>&1
tail -3 gpurun_out/r2_41_pytest.log
cat gpurun_out/r2_41_segce_timing.log
