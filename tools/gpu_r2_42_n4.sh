#!/bin/bash
# round 2, visit 42 (4 GPUs): the bench line at N = 4 on the tree at round end
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/r2_42_bench_n4.out 2> gpurun_out/r2_42_bench_n4.err; echo "exit $?" >> gpurun_out/r2_42_bench_n4.err
tail -3 gpurun_out/r2_42_bench_n4.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_42_bench_n4.out"):
    if l.startswith("{"):
        j = json.loads(l)
        print({k: j[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", j["e2e"]["value"])
        b = j.get("bank") or {}
        print("bank", b.get("ms_per_step"), b.get("value"), b.get("allgather_ms"))
        st = j.get("strong") or {}
        print("strong", st.get("ms_per_step"), st.get("value"))
        sp = j.get("sparse_reset") or {}
        print("sparse", sp.get("ms_per_step"), sp.get("value"))
PY
