#!/bin/bash
# round 2, visit 37 (1 GPU): balanced tile walk of the transposed POS sweep — parity (incl. full sizes), bank step timing
# at configs[2] and configs[3], launch list; initcheck again (packet buffers zero-filled once)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_37_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_37_pytest_gpu.log
timeout 300 python tools/bank_probe.py > gpurun_out/r2_37_bank_probe.log 2>&1
PROBE_MAX_VIEWS=1 timeout 300 python tools/bank_probe.py >> gpurun_out/r2_37_bank_probe.log 2>&1
PROBE_STEPS=3 timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
   --log-file gpurun_out/r2_37_launches_bank.csv python tools/bank_probe.py > gpurun_out/r2_37_ncu_bank.log 2>&1
timeout 600 python bench.py --workload s3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_37_bench_s3.json 2> gpurun_out/r2_37_bench_s3.err; echo "exit $?" >> gpurun_out/r2_37_bench_s3.err
PCL_TC_VARIANT=16 timeout 600 python bench.py --workload s3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_37_bench_s3_rowtile_pos.json 2>> gpurun_out/r2_37_bench_s3.err
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_zpending.py tests/test_gpu_parity.py -m gpu -q -x -k "graphed_bank or many_ranks or shadow_tracks or by_class_blocks" > gpurun_out/r2_37_initcheck.log 2>&1; echo "initcheck exit $?" >> gpurun_out/r2_37_initcheck.log
tail -3 gpurun_out/r2_37_pytest_gpu.log
cat gpurun_out/r2_37_bank_probe.log
python tools/ncu_agg.py gpurun_out/r2_37_launches_bank.csv | grep -E "k_tc|finalize|sum"
python - <<'PY'
import json
for f in ("gpurun_out/r2_37_bench_s3.json", "gpurun_out/r2_37_bench_s3_rowtile_pos.json"):
    for l in open(f):
        if l.startswith("{"):
            j = json.loads(l); print(f, j["ms_per_step"], j["value"])
PY
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/r2_37_initcheck.log | tail -4
