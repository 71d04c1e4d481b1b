#!/bin/bash
# One GPU-box visit: build check, GPU parity tests, smoke, bench, launch list.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print(os.cpu_count())" > gpurun_out/cpus.txt
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 180 python tools/tc_selftest.py dump > gpurun_out/tc_dump.log 2>&1; echo "dump exit $?" >> gpurun_out/tc_dump.log
timeout 300 python tools/tc_selftest.py fwd > gpurun_out/tc_fwd.log 2>&1; echo "fwd exit $?" >> gpurun_out/tc_fwd.log
timeout 300 python tools/tc_selftest.py bwd > gpurun_out/tc_bwd.log 2>&1; echo "bwd exit $?" >> gpurun_out/tc_bwd.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 120 python tools/segce_bench.py > gpurun_out/segce.log 2>&1
timeout 300 python tools/sweep_bench.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep exit $?" >> gpurun_out/sweep.err
if [ "${2:-}" = "variants" ]; then
  for v in 8 9; do
    PCL_TC_VARIANT=$v timeout 120 python tools/sweep_bench.py 16384x65536 65536x131072 1024x190000 >> gpurun_out/sweep_variants.jsonl 2>> gpurun_out/sweep.err
  done
fi
if [ "${1:-}" = "ncutc" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -c 3 -o gpurun_out/tc_prof \
    python tools/sweep_bench.py 65536x131072 > gpurun_out/ncu_tc.log 2>&1
fi
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
cat gpurun_out/tc_dump.log gpurun_out/tc_fwd.log gpurun_out/tc_bwd.log; tail -15 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/segce.log; cat gpurun_out/sweep.jsonl; cat gpurun_out/sweep_variants.jsonl 2>/dev/null | cut -c1-200; tail -2 gpurun_out/sweep.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
