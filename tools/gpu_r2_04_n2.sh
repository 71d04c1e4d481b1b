#!/bin/bash
# 2 GPUs: NCCL bank merge (eager + graphed split step) and the bench line at N=2
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   tools/dist_bank_check.py > gpurun_out/r2_04_dist_bank_check_n2.log 2>&1; echo "exit $?" >> gpurun_out/r2_04_dist_bank_check_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_04_bench_n2.json 2> gpurun_out/r2_04_bench_n2.err; echo "exit $?" >> gpurun_out/r2_04_bench_n2.err
grep -E "OK|FAILED|exit|Error" gpurun_out/r2_04_dist_bank_check_n2.log | tail
cat gpurun_out/r2_04_bench_n2.json; tail -5 gpurun_out/r2_04_bench_n2.err
