#!/bin/bash
# round 2, visit 34 (1 GPU): scan with un-chained loads, selection with vector key loads — timeline + selection tests
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 200 python tools/fused_timeline.py > gpurun_out/r2_34_timeline_full.log 2>&1
PROBE_B=1 timeout 200 python tools/fused_timeline.py > gpurun_out/r2_34_timeline_b1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zpending.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2_34_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2_34_pytest.log
grep -E "replay 2|keys|select|fused  |scatter|fill " gpurun_out/r2_34_timeline_full.log | tail -6
grep -E "replay 2|keys|select|fused  |scatter|fill " gpurun_out/r2_34_timeline_b1.log | tail -6
tail -3 gpurun_out/r2_34_pytest.log
