#!/bin/bash
# round 2, visit 24 (1 GPU): full strict suite on the final tree (last-writer bank apply, branch-free NEGW), smoke, bench
# (train_iter now twice: engine loss step / reference op sequence on the GPU), tensor-path top-k timing, bank probe
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/r2_24_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_24_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_24_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_24_smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_24_bench_s1.json 2> gpurun_out/r2_24_bench_s1.err; echo "exit $?" >> gpurun_out/r2_24_bench_s1.err
PROBE_PATH=tc timeout 300 python tools/topk_probe.py > gpurun_out/r2_24_topk_tc.log 2>&1; echo "exit $?" >> gpurun_out/r2_24_topk_tc.log
timeout 300 python tools/bank_probe.py > gpurun_out/r2_24_bank_probe.log 2>&1; echo "exit $?" >> gpurun_out/r2_24_bank_probe.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/r2_24_pytest_gpu.log | tail -20
tail -3 gpurun_out/r2_24_smoke.log
tail -c 1500 gpurun_out/r2_24_bench_s1.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_24_bench_s1.json"):
    if l.startswith("{"):
        j = json.loads(l)
        print({k: j[k] for k in ("value", "ms_per_step", "gpu_launches")}, j["e2e"]["value"], j["roofline"]["frac"])
        print("bank", j.get("bank", {}).get("ms_per_step"), "sparse", j.get("sparse_reset", {}).get("ms_per_step"))
        print("train", json.dumps(j["config"].get("train_iter"))[:900])
PY
tail -15 gpurun_out/r2_24_topk_tc.log
tail -15 gpurun_out/r2_24_bank_probe.log
