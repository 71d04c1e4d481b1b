"""In-graph timeline of the fused S1 step (diagnostics): per-kernel start/end inside ONE graph replay from %globaltimer
stamps (ncu serialises the kernels of a graph; this shows how they overlap), plus the per-CTA phase stamps of the fused
InfoNCE kernel."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contrastiveseg_b200 as cs
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
cfg = dict(bench.S1)
cfg["B"] = int(os.environ.get("PROBE_B", cfg["B"]))
inp = {k: v.to(dev) for k, v in bench.make_inputs(cfg, 304, None, False).items()}
crit = cs.PixelContrastLoss(bench.engine_configer(cfg, False, "bf16"))
opts = crit.options(); opts.num_classes = cfg["K"]
st = cs.GraphedContrastStep(inp["embed"], inp["target"], seg=inp["seg"], options=opts,
                            sparse_reset=bool(os.environ.get("PROBE_SPARSE")))
assert st.fused
for _ in range(20):
    st.replay()
torch.cuda.synchronize()
sync = st.ws.sync
names = ["keys+plan", "select", "fused", "scatter", "fill"]
I64MAX = (1 << 63) - 1
for rep in range(3):
    tl = sync[16:16 + 2 * 2 * 8].view(torch.int64)
    tl[0::2] = I64MAX
    tl[1::2] = 0
    sync[64:64 + 16 * 32].zero_()
    sync[7] = 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.replay(); e1.record()
    torch.cuda.synchronize()
    sync[7] = 0
    t = tl.cpu().tolist()
    t0 = min(t[2 * k] for k in range(5))
    print(f"--- replay {rep}: event time {e0.elapsed_time(e1) * 1e3:.1f} us")
    for k, n in enumerate(names):
        print(f"  {n:10s} start {(t[2 * k] - t0) / 1e3:7.1f} us   end {(t[2 * k + 1] - t0) / 1e3:7.1f} us   dur {(t[2 * k + 1] - t[2 * k]) / 1e3:7.1f}")
    print(f"    scan part ends (last block) {(t[2 * 5 + 1] - t0) / 1e3:7.1f} us   totals done {(t[2 * 6 + 1] - t0) / 1e3:7.1f} us")
    stamps = sync[64:64 + 16 * 32].view(torch.int64).view(32, 8).cpu()
    live = stamps[:, 0] > 0
    if live.any():
        s = stamps[live]
        f0 = t[2 * 2]
        lab = ["epi start", "MMA1 done", "NEG written", "barrier1", "POS written", "barrier2", "H written", "dA stored"]
        for i, l in enumerate(lab):
            col = (s[:, i] - f0).double() / 1e3
            print(f"    fused {l:12s} min {col.min():6.1f}  max {col.max():6.1f} us (since first CTA start, {int(live.sum())} CTAs)")
