#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
for r in 32 40 0; do echo "== reserve $r SMs"; PCL_FILL_RESERVE_SMS=$r timeout 300 python tools/fused_timeline.py 2>&1 | tail -14; done > gpurun_out/r2_14_timeline_reserve.log
for r in 32 40 48; do echo "== reserve $r"; PCL_FILL_RESERVE_SMS=$r PROBE_FILL=1 PROBE_STEPS=300 timeout 300 python tools/fused_probe.py 2>&1 | head -1; done > gpurun_out/r2_14_probe.log
cat gpurun_out/r2_14_timeline_reserve.log; cat gpurun_out/r2_14_probe.log
