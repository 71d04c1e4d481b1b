#!/bin/bash
# round 2, visit 43 (2 GPUs): all_gather under the backward sweep (third graph for the bank write) — NCCL check, bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   tools/dist_bank_check.py > gpurun_out/r2_43_dist_bank_check_n2.log 2>&1; echo "exit $?" >> gpurun_out/r2_43_dist_bank_check_n2.log
grep -E "OK|FAILED|exit|Error" gpurun_out/r2_43_dist_bank_check_n2.log | tail
PCL_BENCH_NO_TRAIN_ITER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_43_bench_n2.out 2> gpurun_out/r2_43_bench_n2.err; echo "exit $?" >> gpurun_out/r2_43_bench_n2.err
PCL_GATHER_OVERLAP=0 PCL_BENCH_NO_TRAIN_ITER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
   bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_43_bench_n2_no_overlap.out 2>> gpurun_out/r2_43_bench_n2.err
tail -2 gpurun_out/r2_43_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_43_bench_n2.out", "gpurun_out/r2_43_bench_n2_no_overlap.out"):
    for l in open(f):
        if l.startswith("{"):
            j = json.loads(l)
            b = j.get("bank") or {}
            print(f, {k: j[k] for k in ("value", "ms_per_step")}, "bank", b.get("ms_per_step"), b.get("value"), b.get("allgather_ms"))
PY
