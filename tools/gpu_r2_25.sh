#!/bin/bash
# round 2, visit 25 (1 GPU): fresh ncu --set full + source of the bank step's tensor sweeps (after the multiply-shift labels)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
PROBE_STEPS=2 timeout 900 ncu --graph-profiling node --set full --import-source on --clock-control none -k regex:"k_tc_fwd|k_tc_bwd" -c 6 \
   -o gpurun_out/r2_25_bank_tc_full python tools/bank_probe.py > gpurun_out/r2_25_ncu_bank_full.log 2>&1
tail -5 gpurun_out/r2_25_ncu_bank_full.log
ls -la gpurun_out/*.ncu-rep
