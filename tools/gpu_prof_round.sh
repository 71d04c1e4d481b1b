#!/bin/bash
# Profiling visit: (1) launch list of the S1 bench step, (2) ncu --set full of the tensor backward at one S4 point,
# (3) ncu --set full of the S1 step's own kernels (DRAM traffic for the roofline object).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "torch_cpu or trainer_hook" > gpurun_out/pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/pytest_new.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_s1.csv \
  python bench.py --steps 6 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_bwd -c 1 -o gpurun_out/tc_bwd_prof -f \
  python tools/sweep_bench.py 16384x65536 > gpurun_out/ncu_tc_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_scatter|k_keys|k_select|k_plan" -s 40 -c 4 -o gpurun_out/s1_prof -f \
  python bench.py --steps 6 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_s1.log 2>&1
tail -4 gpurun_out/pytest_new.log; tail -2 gpurun_out/ncu_tc_bwd.log; tail -2 gpurun_out/ncu_s1.log
