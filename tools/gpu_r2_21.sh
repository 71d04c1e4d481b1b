#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
for w in 8 16; do echo "== bwd epilogue warps $w"; PCL_TC_BWD_EPI=$w timeout 300 python tools/bank_probe.py; done > gpurun_out/r2_21_bank_probe.log 2>&1
for w in 8 16; do echo "== bwd epilogue warps $w"; PCL_TC_BWD_EPI=$w timeout 300 python tools/sweep_bench.py 1024x190000 16384x65536 | cut -c1-330; done > gpurun_out/r2_21_sweep.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_zpending.py -m gpu -q -x --tb=short > gpurun_out/r2_21_pytest.log 2>&1
cat gpurun_out/r2_21_bank_probe.log gpurun_out/r2_21_sweep.log; tail -3 gpurun_out/r2_21_pytest.log
