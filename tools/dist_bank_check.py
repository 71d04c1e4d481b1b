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
print(f"rank {rank} dist bank check {'OK' if ok else 'FAILED'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
