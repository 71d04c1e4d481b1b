#!/bin/bash
# 2-GPU visit: multi-rank bank merge check + bench at N=2 (s1 weak scaling, s2 bank + allgather).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  tools/dist_bank_check.py > gpurun_out/dist_bank.log 2>&1; echo "dist exit $?" >> gpurun_out/dist_bank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?" >> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --steps 200 --warmup 10 --workload s2 > gpurun_out/bench_s2_n2.json 2> gpurun_out/bench_s2_n2.err; echo "exit $?" >> gpurun_out/bench_s2_n2.err
timeout 300 python bench.py --steps 200 --warmup 10 --workload s2 --no-cpu-baseline > gpurun_out/bench_s2_n1.json 2> gpurun_out/bench_s2_n1.err; echo "exit $?" >> gpurun_out/bench_s2_n1.err
tail -12 gpurun_out/dist_bank.log; cat gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err; cat gpurun_out/bench_s2_n2.json; tail -2 gpurun_out/bench_s2_n2.err; cat gpurun_out/bench_s2_n1.json; tail -2 gpurun_out/bench_s2_n1.err
