#!/bin/bash
# compute-sanitizer memcheck over a subset of the GPU parity tests (small shapes) and the smoke step.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 30 python -m pytest tests -m gpu -q -x \
  -k "loss_and_grad_match_reference or bank_enqueue_matches or tc_forward_matches_oracle or tc_backward_matches_oracle or loss_module_on_tensor_path or fused_upsample or l2_normalize or workspace_sizes" \
  > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed|exit" gpurun_out/sanitize_memcheck.log | tail -15
