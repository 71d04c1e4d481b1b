#!/bin/bash
# round 2, GPU visit 5 (1 GPU): the fused small-anchor step on hardware
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_zpending.py -m gpu -q -rA --tb=short -p no:cacheprovider -x > gpurun_out/r2_05_pytest_zp.log 2>&1; echo "exit $?" >> gpurun_out/r2_05_pytest_zp.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_05_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_05_pytest_gpu.log
PCL_BENCH_HEADLINE_ONLY=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_05_bench_s1.json 2> gpurun_out/r2_05_bench_s1.err; echo "exit $?" >> gpurun_out/r2_05_bench_s1.err
PCL_BENCH_HEADLINE_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_05_launches_s1.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_05_ncu_bench.log 2>&1; echo "ncu exit $?" >> gpurun_out/r2_05_ncu_bench.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r2_05_pytest_zp.log | tail -20
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_05_pytest_gpu.log | tail -10
cat gpurun_out/r2_05_bench_s1.json; tail -3 gpurun_out/r2_05_bench_s1.err
tail -40 gpurun_out/r2_05_launches_s1.csv | cut -c1-220
