#!/bin/bash
# round 2, visit 28 (1 GPU): fill CTAs that own their SMs, forked at the start of the S1 step
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
out=gpurun_out/r2_28_fill_excl.log
: > $out
for g in 24 32 40 48 64; do
  echo "== early fork, PCL_FILL_EXCL=$g" >> $out
  PCL_FILL_FORK_EARLY=1 PCL_FILL_EXCL=$g timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
done
for g in 48 148; do
  echo "== late fork, PCL_FILL_EXCL=$g" >> $out
  PCL_FILL_EXCL=$g timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
done
cat $out
