#!/bin/bash
# round 2, visit 27 (1 GPU): throttled fill forked at the START of the S1 step (rate-limited so that the latency-bound
# chain next to it is not starved) vs the full-rate fill forked after the selection
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
out=gpurun_out/r2_27_fill_throttle.log
: > $out
echo "== baseline (fork after select, 148 CTAs)" >> $out
timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
for g in 12 16 20 24 32 48; do
  echo "== early fork, PCL_FILL_GRID=$g" >> $out
  PCL_FILL_FORK_EARLY=1 PCL_FILL_GRID=$g timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
done
for g in 32 64; do
  echo "== late fork, PCL_FILL_GRID=$g" >> $out
  PCL_FILL_GRID=$g timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
done
cat $out
