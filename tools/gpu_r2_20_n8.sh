#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
   bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_20_bench_n8.json 2> gpurun_out/r2_20_bench_n8.err; echo "exit $?" >> gpurun_out/r2_20_bench_n8.err
grep "^{" gpurun_out/r2_20_bench_n8.json | head -c 6000; tail -5 gpurun_out/r2_20_bench_n8.err
