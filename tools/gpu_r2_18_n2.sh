#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   tools/dist_bank_check.py > gpurun_out/r2_18_dist_bank_check_n2.log 2>&1; echo "exit $?" >> gpurun_out/r2_18_dist_bank_check_n2.log
grep -E "OK|FAILED|exit|Error" gpurun_out/r2_18_dist_bank_check_n2.log | tail
