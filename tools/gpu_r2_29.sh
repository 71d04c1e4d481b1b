#!/bin/bash
# round 2, visit 29 (1 GPU): exclusive fill + max-shared carve-out preference on the chain kernels
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
out=gpurun_out/r2_29_fill_excl_carveout.log
: > $out
for g in 40 48 56 64; do
  echo "== early fork, PCL_FILL_EXCL=$g, PCL_CARVEOUT_MAX=1" >> $out
  PCL_CARVEOUT_MAX=1 PCL_FILL_FORK_EARLY=1 PCL_FILL_EXCL=$g timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
done
echo "== baseline + PCL_CARVEOUT_MAX=1" >> $out
PCL_CARVEOUT_MAX=1 timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
cat $out
