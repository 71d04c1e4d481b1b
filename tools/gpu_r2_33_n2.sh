#!/bin/bash
# round 2, visit 33 (2 GPUs): NCCL check of the bank merge + captured step, and the bench line at N = 2 on the final tree
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   tools/dist_bank_check.py > gpurun_out/r2_33_dist_bank_check_n2.log 2>&1; echo "exit $?" >> gpurun_out/r2_33_dist_bank_check_n2.log
grep -E "OK|FAILED|exit|Error" gpurun_out/r2_33_dist_bank_check_n2.log | tail
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_33_bench_n2.out 2> gpurun_out/r2_33_bench_n2.err; echo "exit $?" >> gpurun_out/r2_33_bench_n2.err
tail -3 gpurun_out/r2_33_bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_33_bench_n2.out"):
    if l.startswith("{"):
        j = json.loads(l)
        print({k: j[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", j["e2e"]["value"])
        print("bank", json.dumps(j.get("bank"))[:400])
        print("strong", json.dumps(j.get("strong"))[:300])
PY
