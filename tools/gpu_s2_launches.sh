#!/bin/bash
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 90 --csv --log-file gpurun_out/launches_s2.csv \
  python bench.py --workload s2 --steps 6 --warmup 6 --no-cpu-baseline > gpurun_out/ncu_s2.log 2>&1
tail -2 gpurun_out/ncu_s2.log | cut -c1-300
