#!/bin/bash
# round 2, visit 31 (1 GPU): late fork (after the selection) with fill CTAs that own 148 - 32 SMs: the fused InfoNCE
# kernel gets its 32 SMs at once instead of waiting for fill CTAs to retire
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
out=gpurun_out/r2_31_fill_excl_late.log
: > $out
for g in 96 108 116; do
  echo "== late fork, PCL_FILL_EXCL=$g" >> $out
  PCL_FILL_EXCL=$g timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
done
cat $out
