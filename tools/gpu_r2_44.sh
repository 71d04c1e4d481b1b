#!/bin/bash
# round 2, visit 40 (1 GPU): confirmation of the tree at round end — full suite, timeline, smoke, bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_44_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2_44_pytest_gpu.log
timeout 200 python tools/fused_timeline.py > gpurun_out/r2_44_timeline_full.log 2>&1
PROBE_B=1 timeout 200 python tools/fused_timeline.py > gpurun_out/r2_44_timeline_b1.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_44_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2_44_smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_44_bench_s1.json 2> gpurun_out/r2_44_bench_s1.err; echo "exit $?" >> gpurun_out/r2_44_bench_s1.err
tail -3 gpurun_out/r2_44_pytest_gpu.log
grep -E "replay 2|keys|select|fused  |scatter|fill " gpurun_out/r2_44_timeline_full.log | tail -6
grep -E "replay 2|keys|select|fused  |scatter|fill " gpurun_out/r2_44_timeline_b1.log | tail -6
tail -2 gpurun_out/r2_44_smoke.log
python - <<'PY'
import json
for l in open("gpurun_out/r2_44_bench_s1.json"):
    if l.startswith("{"):
        j = json.loads(l)
        print({k: j[k] for k in ("value", "ms_per_step", "gpu_launches")}, j["e2e"]["value"], json.dumps(j["roofline"])[:600])
        print("bank", j.get("bank", {}).get("ms_per_step"), "sparse", j.get("sparse_reset", {}).get("ms_per_step"))
PY
