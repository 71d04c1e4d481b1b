#!/bin/bash
# Evidence captures for profiles/: tests, bench, launch list, ncu --set full of the NEG sweep and of the S1 writer.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --steps 200 --warmup 10 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 100 --csv --log-file gpurun_out/launches_s1.csv \
  python bench.py --steps 6 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_fwd -c 2 -o gpurun_out/tc_fwd_prof -f \
  python tools/sweep_bench.py 65536x131072 > gpurun_out/ncu_tc_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_zero_scatter|k_keys" -s 20 -c 2 -o gpurun_out/s1_prof -f \
  python bench.py --steps 6 --warmup 8 --no-cpu-baseline > gpurun_out/ncu_s1.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-900; cat gpurun_out/bench_fp32.json | cut -c1-300; tail -2 gpurun_out/bench.err
