#!/bin/bash
# ncu --set full on the tensor sweep kernels (fwd NEG, fwd POS, bwd) at one S4 point.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -c 3 -o gpurun_out/tc_prof -f \
  python tools/sweep_bench.py 32768x65536 > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
