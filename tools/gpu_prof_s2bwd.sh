#!/bin/bash
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_bwd -s 8 -c 1 -o gpurun_out/s2_bwd_prof -f \
  python bench.py --workload s2 --steps 6 --warmup 6 --no-cpu-baseline > gpurun_out/ncu_s2bwd.log 2>&1
tail -2 gpurun_out/ncu_s2bwd.log | cut -c1-200
