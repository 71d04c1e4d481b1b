#!/bin/bash
# round 2, GPU visit 3 (1 GPU): strict full suite after the smem fix, new full-size / aux / graphed-bank tests, smoke, bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/r2_03_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_03_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_03_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_03_smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_03_bench_s1.json 2> gpurun_out/r2_03_bench_s1.err; echo "exit $?" >> gpurun_out/r2_03_bench_s1.err
timeout 600 python bench.py --workload s3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_03_bench_s3.json 2> gpurun_out/r2_03_bench_s3.err; echo "exit $?" >> gpurun_out/r2_03_bench_s3.err
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_03_pytest_gpu.log | tail -20
tail -5 gpurun_out/r2_03_smoke.log
cat gpurun_out/r2_03_bench_s1.json; tail -3 gpurun_out/r2_03_bench_s1.err
cat gpurun_out/r2_03_bench_s3.json; tail -3 gpurun_out/r2_03_bench_s3.err
