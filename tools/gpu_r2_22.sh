#!/bin/bash
# a10 evidence (ncu of the top-k sweeps) + sanitizer passes over the round-2 kernels
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 300 python tools/topk_probe.py > gpurun_out/r2_22_topk_probe.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_topk" -c 12 -o gpurun_out/r2_22_topk_full python tools/topk_probe.py > gpurun_out/r2_22_ncu_topk.log 2>&1
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_zpending.py -m gpu -q -x -k "fused or graphed_bank or sparse" > gpurun_out/r2_22_memcheck_fused.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_22_memcheck_fused.log
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_zpending.py -m gpu -q -x -k "fused_small_anchor and geom0 or sparse" > gpurun_out/r2_22_racecheck_fused.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_22_racecheck_fused.log
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_zpending.py -m gpu -q -x -k "fused_small_anchor and geom1" > gpurun_out/r2_22_initcheck_fused.log 2>&1; echo "initcheck exit $?" >> gpurun_out/r2_22_initcheck_fused.log
cat gpurun_out/r2_22_topk_probe.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/r2_22_memcheck_fused.log gpurun_out/r2_22_racecheck_fused.log gpurun_out/r2_22_initcheck_fused.log | tail -12
