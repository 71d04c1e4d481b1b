#!/bin/bash
# round 2, visit 36 (1 GPU): ncu --set full of the new kernels (DRAM traffic for the roofline line), sanitizer passes over them
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
PROBE_FILL=1 PROBE_STEPS=2 timeout 900 ncu --graph-profiling node --set full --import-source on --clock-control none -k regex:"k_fill_zero_excl|k_self_fused|k_keys|k_select|k_scatter_reduce" -c 10 \
   -o gpurun_out/r2_36_s1_step_full python tools/fused_probe.py > gpurun_out/r2_36_ncu_s1.log 2>&1
PROBE_STEPS=2 timeout 900 ncu --graph-profiling node --set full --import-source on --clock-control none -k regex:"k_tc_pos_t|k_bank_apply" -c 4 \
   -o gpurun_out/r2_36_bank_pos_full python tools/bank_probe.py > gpurun_out/r2_36_ncu_bank.log 2>&1
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_zpending.py tests/test_gpu_parity.py -m gpu -q -x -k "fused_small_anchor or graphed_bank or many_ranks or shadow_tracks or trainer_hook_end_to_end" > gpurun_out/r2_36_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_36_memcheck.log
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_zpending.py tests/test_gpu_parity.py -m gpu -q -x -k "graphed_bank or many_ranks or shadow_tracks" > gpurun_out/r2_36_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_36_racecheck.log
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_zpending.py tests/test_gpu_parity.py -m gpu -q -x -k "graphed_bank or many_ranks or shadow_tracks" > gpurun_out/r2_36_initcheck.log 2>&1; echo "initcheck exit $?" >> gpurun_out/r2_36_initcheck.log
tail -3 gpurun_out/r2_36_ncu_s1.log; tail -3 gpurun_out/r2_36_ncu_bank.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/r2_36_memcheck.log gpurun_out/r2_36_racecheck.log gpurun_out/r2_36_initcheck.log | tail -12
