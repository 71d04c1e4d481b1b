#!/bin/bash
# round 2, visit 26 (1 GPU): transposed POS sweep of the bank mode (k_tc_pos_t) — parity, timing, launch list; apply probe
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_zpending.py tests/test_gpu_topk.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2_26_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2_26_pytest.log
timeout 300 python tools/bank_probe.py > gpurun_out/r2_26_bank_probe.log 2>&1
PROBE_MAX_VIEWS=1 timeout 300 python tools/bank_probe.py >> gpurun_out/r2_26_bank_probe.log 2>&1
PCL_TC_VARIANT=16 timeout 300 python tools/bank_probe.py >> gpurun_out/r2_26_bank_probe.log 2>&1
PROBE_STEPS=3 timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
   --log-file gpurun_out/r2_26_launches_bank.csv python tools/bank_probe.py > gpurun_out/r2_26_ncu_bank.log 2>&1
timeout 300 python tools/apply_probe.py > gpurun_out/r2_26_apply_probe.log 2>&1
tail -5 gpurun_out/r2_26_pytest.log
cat gpurun_out/r2_26_bank_probe.log
python tools/ncu_agg.py gpurun_out/r2_26_launches_bank.csv
cat gpurun_out/r2_26_apply_probe.log | tail -8
