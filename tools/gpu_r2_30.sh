#!/bin/bash
# round 2, visit 30 (1 GPU): fill CTAs that take ALL shared memory of their SM (true exclusivity), early fork
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
out=gpurun_out/r2_30_fill_excl_full_smem.log
: > $out
for g in 32 40 48 56 64; do
  echo "== early fork, PCL_FILL_EXCL=$g" >> $out
  PCL_FILL_FORK_EARLY=1 PCL_FILL_EXCL=$g timeout 200 python tools/fused_timeline.py 2>&1 | grep -E "replay 2|keys|select|fused  |scatter|fill " | tail -6 >> $out
done
cat $out
