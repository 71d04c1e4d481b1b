#!/bin/bash
# bench line with the fused step + ncu --set full of the step's kernels (traffic) + launch list
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m contrastiveseg_b200.build > gpurun_out/build.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r2_15_bench_s1.json 2> gpurun_out/r2_15_bench_s1.err; echo "exit $?" >> gpurun_out/r2_15_bench_s1.err
PROBE_STEPS=2 PROBE_FILL=1 timeout 900 ncu --graph-profiling node --set full --import-source on --clock-control none \
   -k regex:"k_fill_zero|k_keys|k_select|k_self_fused|k_scatter_reduce" -c 12 -o gpurun_out/r2_15_s1_step_full python tools/fused_probe.py > gpurun_out/r2_15_ncu_full.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_15_smoke.log 2>&1
cat gpurun_out/r2_15_bench_s1.json; tail -3 gpurun_out/r2_15_bench_s1.err; tail -3 gpurun_out/r2_15_smoke.log; tail -3 gpurun_out/r2_15_ncu_full.log
