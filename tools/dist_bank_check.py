"""torchrun --nproc-per-node N tools/dist_bank_check.py : multi-GPU bank merge check (NCCL allgather of the
enqueue packets).  Every rank enqueues its own image; afterwards all banks must be bit-identical and equal to a
single process enqueueing rank 0's image, then rank 1's, ... (oracle port, replayed permutations)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.distributed as dist

import contrastiveseg_b200 as cs
from contrastiveseg_b200.synth import make_bank, make_contrast_batch
from oracle import ref_port as P

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
K, D, M, F, steps = 19, 256, 50, 10, 4
bank0 = make_bank(K, M, D, 11)
names = ("segment_queue", "segment_queue_ptr", "pixel_queue", "pixel_queue_ptr")
mine = [bank0[k].clone().to(dev) for k in names]
ref = [bank0[k].clone() for k in names]
for s in range(steps):
    datas = [make_contrast_batch(B=1, D=D, h=32, w=64, num_classes=K, img_stride=4, block=16, seed=1000 * s + r)
             for r in range(world)]
    perms = []
    for r in range(world):          # same recorded permutations on both sides (per rank, reference call order)
        rec = P.PermRecorder(torch.Generator().manual_seed(77 * s + r))
        P.dequeue_and_enqueue(datas[r]["embed"], datas[r]["target"], *ref, network_stride=4, memory_size=M,
                              pixel_update_freq=F, perm_fn=rec)
        perms.append(rec.draws)
    cs.dequeue_and_enqueue(datas[rank]["embed"].to(dev), datas[rank]["target"].to(dev), *mine, network_stride=4,
                           memory_size=M, pixel_update_freq=F, perm_fn=P.PermReplay(perms[rank]))
torch.cuda.synchronize()
ok = True
for name, a, b in zip(names, mine, ref):
    d = (a.cpu().double() - b.double()).abs().max().item()
    lim = 0 if "ptr" in name else 2e-6
    print(f"rank {rank} {name}: max diff vs sequential oracle {d:.2e}", flush=True)
    ok &= d <= lim
# bit-identical across ranks
for a in mine:
    g = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(g, a)
    ok &= all(torch.equal(g[0], x) for x in g)
# ---- the graphed bank step on several ranks: two graphs around ONE all_gather of the packet (graph_step.py).  Replay r
# must equal the eager trainer order (loss -> enqueue [all_gather] -> backward -> deferred write) on every rank, and the
# banks must stay bit-identical across ranks. ----
from contrastiveseg_b200 import bank as bank_mod, functional as Fn
K2, M2 = 7, 48
data = make_contrast_batch(B=1, D=256, h=32, w=32, num_classes=K2, img_stride=4, block=16, seed=500 + rank)
embed, tgt, seg = data["embed"].to(dev), data["target"].to(dev), data["seg"].to(dev)
torch.manual_seed(0)
bank_g = cs.MemoryBank(K2, M2, 256, with_shadow=True).to(dev)
bank_e = cs.MemoryBank(K2, M2, 256, with_shadow=True).to(dev)
bank_e.load_state_dict(bank_g.state_dict())
bank_g.sync_shadow(); bank_e.sync_shadow()
opts = cs.ContrastOptions(temperature=0.07, base_temperature=0.07, max_samples=128, max_views=8, seed=5, precision="bf16",
                          num_classes=K2)
step = cs.GraphedContrastStep(embed, tgt, seg=seg, segment_queue=bank_g.segment_queue, pixel_queue=bank_g.pixel_queue,
                              bank_shadow=bank_g.shadow, options=opts,
                              enqueue=dict(bank=bank_g, network_stride=4, pixel_update_freq=5, seed=3))
gok = step.split and step.graph_b is not None
for r in range(3):
    loss, grad = step.replay()
    torch.cuda.synchronize()
    Fn._step_counter[0] = r
    bank_mod._enqueue_counter[0] = r
    e = embed.clone().requires_grad_(True)
    l = cs.pixel_contrast_loss(e, tgt, seg=seg, segment_queue=bank_e.segment_queue, pixel_queue=bank_e.pixel_queue,
                               bank_shadow=bank_e.shadow, options=opts)
    bank_e.enqueue(e.detach(), tgt, network_stride=4, pixel_update_freq=5, seed=3)
    l.backward()
    torch.cuda.synchronize()
    gok &= torch.equal(l.detach(), loss) and torch.allclose(e.grad, grad, rtol=2e-6, atol=0)
    for name in names:
        gok &= torch.equal(getattr(bank_g, name), getattr(bank_e, name))
    gok &= torch.equal(bank_g.shadow, bank_e.shadow)
# back-to-back replays (no host synchronisation between the steps: the all_gather of step r runs on its own stream under
# the backward sweep, the bank write and the next step's packet must still be ordered around it)
import os
for r in range(3, 9):
    step.replay()
torch.cuda.synchronize()
for r in range(3, 9):
    Fn._step_counter[0] = r
    bank_mod._enqueue_counter[0] = r
    e = embed.clone().requires_grad_(True)
    l = cs.pixel_contrast_loss(e, tgt, seg=seg, segment_queue=bank_e.segment_queue, pixel_queue=bank_e.pixel_queue,
                               bank_shadow=bank_e.shadow, options=opts)
    bank_e.enqueue(e.detach(), tgt, network_stride=4, pixel_update_freq=5, seed=3)
    l.backward()
torch.cuda.synchronize()
gok &= torch.equal(l.detach(), step.loss) and torch.allclose(e.grad, step.grad, rtol=2e-6, atol=0)
for name in names:
    gok &= torch.equal(getattr(bank_g, name), getattr(bank_e, name))
gok &= torch.equal(bank_g.shadow, bank_e.shadow)
overlap = os.environ.get("PCL_GATHER_OVERLAP", "1") != "0"
gok &= (step.graph_c is not None) == overlap
for name in names:
    a = getattr(bank_g, name)
    g = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(g, a)
    gok &= all(torch.equal(g[0], x) for x in g)
print(f"rank {rank} graphed bank step ({'3 graphs, all_gather under the backward' if overlap else '2 graphs + 1 all_gather'}; 9 steps)", flush=True)
print(f"rank {rank} graphed bank step (2 graphs + 1 all_gather) {'OK' if gok else 'FAILED'}", flush=True)
ok &= bool(gok)
print(f"rank {rank} dist bank check {'OK' if ok else 'FAILED'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
