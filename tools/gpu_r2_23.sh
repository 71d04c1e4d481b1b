#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_23_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_23_pytest_gpu.log
timeout 300 python tools/topk_probe.py > gpurun_out/r2_23_topk_probe.log 2>&1
PROBE_PATH=tc timeout 300 python tools/topk_probe.py >> gpurun_out/r2_23_topk_probe.log 2>&1
PROBE_PATH=tc PROBE_K=64 timeout 300 python tools/topk_probe.py >> gpurun_out/r2_23_topk_probe.log 2>&1
PROBE_PATH=tc timeout 900 ncu --set full --clock-control none -k regex:"k_tc_fwd|k_topk_scan" -c 9 -o gpurun_out/r2_23_tc_topk_full python tools/topk_probe.py > gpurun_out/r2_23_ncu.log 2>&1
PROBE_PATH=tc timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_23_tc_topk_launches.csv python tools/topk_probe.py > /dev/null 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_23_pytest_gpu.log | tail -8
cat gpurun_out/r2_23_topk_probe.log
python tools/ncu_agg.py gpurun_out/r2_23_tc_topk_launches.csv
