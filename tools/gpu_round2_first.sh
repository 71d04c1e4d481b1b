#!/bin/bash
# First GPU-box visit of the next round: everything that was written after the round-1 GPU budget ran out.
#   1. the verified suite (must stay green), 2. the gated a10 top-k parity tests, 3. memcheck over the top-k kernels,
#   4. bench N=1 (per-rank diagnostics need N>1: run tools/gpu_scale8.sh with --gpus 8 separately).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
PCL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_zpending.py -m gpu -q > gpurun_out/pytest_topk.log 2>&1
echo "experimental exit $?" >> gpurun_out/pytest_topk.log
PCL_TEST_EXPERIMENTAL=1 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 \
  --print-limit 30 python -m pytest tests/test_gpu_topk.py -m gpu -q -x -k "exact_data or zero_tail" \
  > gpurun_out/sanitize_topk.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_topk.log
# the CPU emulator executes blocks sequentially: shared-memory races and barrier divergence of the NEW kernels can only
# show up on the hardware -> racecheck + synccheck over the top-k / rank-draw / scatter-only paths (small shapes)
PCL_TEST_EXPERIMENTAL=1 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 \
  --print-limit 20 python -m pytest tests/test_gpu_topk.py tests/test_gpu_zpending.py -m gpu -q -x -k "exact_data or zero_tail or fp32" \
  > gpurun_out/racecheck_new.log 2>&1; echo "racecheck exit $?" >> gpurun_out/racecheck_new.log
PCL_TEST_EXPERIMENTAL=1 PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 \
  --print-limit 20 python -m pytest tests/test_gpu_topk.py -m gpu -q -x -k "exact_data" \
  > gpurun_out/synccheck_new.log 2>&1; echo "synccheck exit $?" >> gpurun_out/synccheck_new.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --steps 200 --warmup 10 --graph --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench --graph exit $?" >> gpurun_out/bench_graph.err
timeout 600 python bench.py --workload s3 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; echo "bench s3 exit $?" >> gpurun_out/bench_s3.err
tail -15 gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_topk.log
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed|exit" gpurun_out/sanitize_topk.log | tail -10
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard|passed|failed|exit" gpurun_out/racecheck_new.log gpurun_out/synccheck_new.log | tail -8
tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_graph.json; tail -3 gpurun_out/bench_graph.err; cat gpurun_out/bench_s3.json; tail -3 gpurun_out/bench_s3.err
