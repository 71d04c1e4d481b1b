#!/bin/bash
# round 2, visit 38-39 (1 GPU): POS sweep ILP; chunk labels from one division per tile — timing and parity
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_zpending.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2_39_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2_39_pytest.log
timeout 300 python tools/bank_probe.py > gpurun_out/r2_39_bank_probe.log 2>&1
PROBE_MAX_VIEWS=1 timeout 300 python tools/bank_probe.py >> gpurun_out/r2_39_bank_probe.log 2>&1
PROBE_STEPS=3 timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
   --log-file gpurun_out/r2_39_launches_bank.csv python tools/bank_probe.py > gpurun_out/r2_39_ncu_bank.log 2>&1
timeout 600 python bench.py --workload s3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_39_bench_s3.json 2> gpurun_out/r2_39_bench_s3.err; echo "exit $?" >> gpurun_out/r2_39_bench_s3.err
tail -3 gpurun_out/r2_39_pytest.log
cat gpurun_out/r2_39_bank_probe.log
python tools/ncu_agg.py gpurun_out/r2_39_launches_bank.csv | grep -E "k_tc|finalize|sum"
python - <<'PY'
import json
for f in ("gpurun_out/r2_39_bench_s3.json",):
    for l in open(f):
        if l.startswith("{"):
            j = json.loads(l); print(f, j["ms_per_step"], j["value"])
PY
