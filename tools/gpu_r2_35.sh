#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 200 python tools/fused_timeline.py > gpurun_out/r2_35_timeline_full.log 2>&1
PROBE_B=1 timeout 200 python tools/fused_timeline.py > gpurun_out/r2_35_timeline_b1.log 2>&1
grep -E "replay 2|keys|scan part|select " gpurun_out/r2_35_timeline_full.log | tail -4
grep -E "replay 2|keys|scan part|select " gpurun_out/r2_35_timeline_b1.log | tail -4
