#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 300 python tools/fused_probe.py > gpurun_out/r2_06_fill_occupancy.log 2>&1
PROBE_B=1 timeout 300 python tools/fused_probe.py >> gpurun_out/r2_06_fill_occupancy.log 2>&1
PROBE_STEPS=3 PROBE_FILL=2 timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
   --log-file gpurun_out/r2_06_launches_fused.csv python tools/fused_probe.py > gpurun_out/r2_06_ncu.log 2>&1
cat gpurun_out/r2_06_fill_occupancy.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_06_launches_fused.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][-60:]
    agg.setdefault(name, []).append(float(r[-1]))
for k, v in agg.items():
    v2 = sorted(v)
    print(f"{k:60s} n={len(v):3d} median={v2[len(v2)//2]/1e3:8.2f} us  min={v2[0]/1e3:8.2f}")
PY
