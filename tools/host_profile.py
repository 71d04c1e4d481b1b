"""cProfile of the host side of the S1 step loop (where do the ~0.14 ms of enqueue time per step go?)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import bench
import contrastiveseg_b200 as cs
dev = torch.device("cuda:0")
cfg = dict(bench.S1)
inp = {k: v.to(dev) for k, v in bench.make_inputs(cfg, 304).items()}
crit = cs.PixelContrastLoss(bench.engine_configer(cfg, False, "bf16"))
embed = inp["embed"].clone().requires_grad_(True)

def step():
    embed.grad = None
    loss = crit(embed, inp["target"], seg=inp["seg"])
    loss.backward()

for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("enqueue us/step", (t1 - t0) / 300 * 1e6, " total us/step", (time.perf_counter() - t0) / 300 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
