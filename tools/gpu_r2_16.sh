#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_19_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_19_pytest_gpu.log
timeout 300 python tools/bank_probe.py > gpurun_out/r2_19_bank_probe.log 2>&1
PROBE_MAX_VIEWS=1 timeout 300 python tools/bank_probe.py >> gpurun_out/r2_19_bank_probe.log 2>&1
PROBE_STEPS=3 timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
   --log-file gpurun_out/r2_19_launches_bank.csv python tools/bank_probe.py > gpurun_out/r2_19_ncu_bank.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_19_pytest_gpu.log | tail -8
cat gpurun_out/r2_19_bank_probe.log
python tools/ncu_agg.py gpurun_out/r2_19_launches_bank.csv
