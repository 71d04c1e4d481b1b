#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 300 python tools/fused_probe.py > gpurun_out/r2_06_fill_occupancy.log 2>&1
PROBE_B=1 timeout 300 python tools/fused_probe.py >> gpurun_out/r2_06_fill_occupancy.log 2>&1
PROBE_STEPS=3 PROBE_FILL=2 timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
   --log-file gpurun_out/r2_06_launches_fused.csv python tools/fused_probe.py > gpurun_out/r2_06_ncu.log 2>&1
timeout 300 python tools/bank_probe.py > gpurun_out/r2_06_bank_probe.log 2>&1
PROBE_MAX_VIEWS=1 timeout 300 python tools/bank_probe.py >> gpurun_out/r2_06_bank_probe.log 2>&1
PROBE_STEPS=3 timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
   --log-file gpurun_out/r2_06_launches_bank.csv python tools/bank_probe.py > gpurun_out/r2_06_ncu_bank.log 2>&1
PROBE_STEPS=2 timeout 900 ncu --graph-profiling node --set full --import-source on --clock-control none -k regex:"k_tc_fwd|k_tc_bwd" -c 6 \
   -o gpurun_out/r2_06_bank_tc_full python tools/bank_probe.py > gpurun_out/r2_06_ncu_bank_full.log 2>&1
timeout 600 python -m pytest tests/test_gpu_topk.py -m gpu -q -k inside_the_captured > gpurun_out/r2_06_pytest_topk_graph.log 2>&1
cat gpurun_out/r2_06_fill_occupancy.log
python tools/ncu_agg.py gpurun_out/r2_06_launches_fused.csv
cat gpurun_out/r2_06_bank_probe.log
python tools/ncu_agg.py gpurun_out/r2_06_launches_bank.csv
tail -3 gpurun_out/r2_06_pytest_topk_graph.log
