#!/bin/bash
# N-GPU runs of the headline workload (gpurun --gpus N): eager weak scaling (the driver's own measurement), the same
# with the step as one CUDA graph (host cost per step -> one launch), and strong scaling (global batch 8 split over ranks).
set -u
N=${1:-8}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
run() {   # name, extra bench flags
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 200 --warmup 10 $2 > gpurun_out/bench_n${N}_$1.json 2> gpurun_out/bench_n${N}_$1.err
  echo "exit $?" >> gpurun_out/bench_n${N}_$1.err
  grep -h "^{" gpurun_out/bench_n${N}_$1.json | cut -c1-400; tail -2 gpurun_out/bench_n${N}_$1.err
}
run weak ""
run weak_graph "--graph"
run strong "--scaling strong"
