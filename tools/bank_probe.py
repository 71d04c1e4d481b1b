"""Replays of the captured memory-bank step (BASELINE configs[2] shape per rank) for ncu / timing runs."""
import json
import os
import sys
import types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
steps = int(os.environ.get("PROBE_STEPS", "100"))
args = types.SimpleNamespace(steps=steps, warmup=5, precision="bf16", workload="s2", scaling="weak", no_graph=False,
                             no_cpu_baseline=True, gpus=1)
cfg = dict(bench.S2)
if os.environ.get("PROBE_MAX_VIEWS"):
    cfg["max_views"] = int(os.environ["PROBE_MAX_VIEWS"])
w = bench.Workload(args, "s2", cfg, True, 0, 1, dev)
assert w.steps, w.graph_error
for i in range(5):
    w.replay(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    w.replay(i)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"workload": "s2", "max_views": cfg["max_views"], "ms_per_step": e0.elapsed_time(e1) / steps,
                  "anchors": int(w.steps[0].ws.plan[2].item()), "kernels_per_step": w.kernels_per_step()}), flush=True)
