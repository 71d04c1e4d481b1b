"""world_size-2 gloo tests (CPU) of the bank merge plumbing: one all_gather of the fixed-size enqueue packet,
rank-major order, identical result on every rank; world 1 is a pass-through.  The packet build / application are
CUDA kernels (GPU: tests/test_gpu_parity.py + tools/dist_bank_check.py over NCCL); the last test here runs their
sources on the host-fiber emulator (tests/emu) on both ranks with the real gloo all_gather in between."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from contrastiveseg_b200.bank import gather_packets


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 37
    packet = torch.arange(n, dtype=torch.float32) + 1000.0 * rank
    gathered = gather_packets(packet)
    assert gathered.shape == (world, n)
    for r in range(world):
        assert torch.equal(gathered[r], torch.arange(n, dtype=torch.float32) + 1000.0 * r)
    torch.save(gathered, os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_packets_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)                      # every rank applies the same packets in the same order


def test_gather_packets_without_process_group_is_passthrough():
    p = torch.arange(5, dtype=torch.float32)
    g = gather_packets(p)
    assert g.shape == (1, 5) and torch.equal(g[0], p)


def _merge_worker(rank, world, port, out_dir):
    """Oracle-level check of the merge semantics: each rank owns different images; applying the gathered packets in
    rank order equals one process enqueueing rank 0's images then rank 1's (the documented Q9 deviation)."""
    from oracle import ref_port as P
    from contrastiveseg_b200.synth import make_bank, make_contrast_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, D, M, F = 5, 16, 6, 3
    data = make_contrast_batch(B=1, D=D, h=12, w=12, num_classes=K, img_stride=2, block=6, seed=100 + rank)
    # the "packet" here is the raw (keys, labels) pair: gather it and replay sequentially on every rank
    flat = torch.cat((data["embed"].reshape(-1), data["target"].reshape(-1).float()))
    allp = gather_packets(flat)
    bank = make_bank(K, M, D, 7)
    bufs = [bank[k] for k in ("segment_queue", "segment_queue_ptr", "pixel_queue", "pixel_queue_ptr")]
    ne = data["embed"].numel()
    for r in range(world):
        keys = allp[r, :ne].reshape(data["embed"].shape)
        labels = allp[r, ne:].reshape(data["target"].shape).long()
        P.dequeue_and_enqueue(keys, labels, *bufs, network_stride=2, memory_size=M, pixel_update_freq=F,
                              perm_fn=lambda n: torch.arange(n))
    torch.save([b.clone() for b in bufs], os.path.join(out_dir, f"b{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_ordered_merge_gives_identical_banks(tmp_path):
    port = _free_port()
    mp.spawn(_merge_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = torch.load(tmp_path / "b0.pt"), torch.load(tmp_path / "b1.pt")
    for x, y in zip(b0, b1):
        assert torch.equal(x, y)


def _emu_bank_worker(rank, world, port, out_dir):
    """The real enqueue kernels (host-fiber emulation of csrc/pcl_bank.cu) on every rank + the real all_gather (gloo):
    what tools/dist_bank_check.py checks on GPUs with NCCL."""
    import pytest
    import emu_harness
    import contrastiveseg_b200 as cs
    from oracle import ref_port as P
    from contrastiveseg_b200.synth import make_bank, make_contrast_batch
    emu_harness.use_emulation(pytest.MonkeyPatch())
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, D, M, F, steps = 6, 32, 9, 4, 3
    names = ("segment_queue", "segment_queue_ptr", "pixel_queue", "pixel_queue_ptr")
    bank0 = make_bank(K, M, D, 11)
    mine = [bank0[k].clone() for k in names]
    ref = [bank0[k].clone() for k in names]
    for s in range(steps):
        datas = [make_contrast_batch(B=2, D=D, h=12, w=16, num_classes=K, img_stride=2, block=4, seed=1000 * s + r)
                 for r in range(world)]
        perms = []
        for r in range(world):                  # sequential oracle: rank 0's images, then rank 1's
            rec = P.PermRecorder(torch.Generator().manual_seed(77 * s + r))
            P.dequeue_and_enqueue(datas[r]["embed"], datas[r]["target"], *ref, network_stride=2, memory_size=M,
                                  pixel_update_freq=F, perm_fn=rec)
            perms.append(rec.draws)
        cs.dequeue_and_enqueue(datas[rank]["embed"].clone(), datas[rank]["target"], *mine, network_stride=2,
                               memory_size=M, pixel_update_freq=F, perm_fn=P.PermReplay(perms[rank]))
    assert torch.equal(mine[1], ref[1]) and torch.equal(mine[3], ref[3])
    assert (mine[0] - ref[0]).abs().max().item() <= 2e-6 and (mine[2] - ref[2]).abs().max().item() <= 2e-7
    torch.save(mine, os.path.join(out_dir, f"e{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_emulated_enqueue_kernels_merge_across_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_emu_bank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = torch.load(tmp_path / "e0.pt"), torch.load(tmp_path / "e1.pt")
    for x, y in zip(b0, b1):
        assert torch.equal(x, y)                    # bit-identical banks on both ranks


def _ddp_hook_worker(rank, world, port, out_dir):
    """C3-style plumbing on CPU: a tiny DDP-wrapped producer ({'seg','embed','key','lb_key'} contract), the drop-in hook
    (MemContrastCELoss + bank enqueue) on the emulated kernels, SGD.  Two warm-up iterations (contrast weighted 0 but the
    projection head stays in the graph, loss_contrast.py:189) and two contrast iterations."""
    import pytest
    import torch.nn as nn
    import emu_harness
    import contrastiveseg_b200 as cs
    from contrastiveseg_b200.synth import make_contrast_batch
    emu_harness.use_emulation(pytest.MonkeyPatch())
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, D = 5, 32

    class TinyNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = nn.Sequential(nn.Conv2d(3, 8, 3, stride=2, padding=1), nn.ReLU())
            self.cls = nn.Conv2d(8, K, 1)
            self.proj = cs.ProjectionHead(8, D, proj="convmlp", bn_type="torchbn")

        def forward(self, x, targets):
            f = self.body(x)
            emb = self.proj(f)
            return {"seg": self.cls(f), "embed": emb, "key": emb.detach(), "lb_key": targets}

    torch.manual_seed(0)                                   # identical initial weights on both ranks
    net = nn.parallel.DistributedDataParallel(TinyNet(), find_unused_parameters=True, broadcast_buffers=False)
    cfg = cs.Configer({"data": {"num_classes": K}, "network": {"stride": 2},
                       "loss": {"loss_type": "mem_contrast_ce_loss", "params": {"ce_ignore_index": -1}},
                       "contrast": {"temperature": 0.1, "base_temperature": 0.07, "max_samples": 48, "max_views": 4,
                                    "loss_weight": 0.1, "use_rmi": False, "use_lovasz": False, "warmup_iters": 2,
                                    "with_memory": True, "memory_size": 12, "pixel_update_freq": 3}})
    torch.manual_seed(1)
    bank = cs.MemoryBank(K, 12, D)                         # same initial bank on both ranks
    hook = cs.ContrastTrainerHook(cfg, bank)
    opt = torch.optim.SGD(net.parameters(), lr=0.05)
    losses = []
    for it in range(4):
        data = make_contrast_batch(B=2, D=D, h=12, w=12, num_classes=K, img_stride=2, block=4, seed=50 * it + rank)
        x = torch.randn(2, 3, 24, 24, generator=torch.Generator().manual_seed(7 * it + rank))
        out = net(x, data["target"])
        loss = hook.loss_step(out, data["target"], iters=it)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(l == l and abs(l) < 1e4 for l in losses)
    proj_w = net.module.proj.proj[0].weight
    assert proj_w.grad is not None                          # the head received a (possibly zero-weighted) gradient
    state = {"params": [p.detach().clone() for p in net.parameters()],
             "bank": [b.clone() for b in (bank.segment_queue, bank.segment_queue_ptr, bank.pixel_queue, bank.pixel_queue_ptr)]}
    torch.save(state, os.path.join(out_dir, f"d{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_training_iterations_with_the_hook_on_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_ddp_hook_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    s0, s1 = torch.load(tmp_path / "d0.pt"), torch.load(tmp_path / "d1.pt")
    for a, b in zip(s0["params"], s1["params"]):
        assert torch.equal(a, b)                            # DDP kept the replicas in step through the custom backward
    for a, b in zip(s0["bank"], s1["bank"]):
        assert torch.equal(a, b)                            # one all_gather per step kept the banks identical
    assert not torch.equal(s0["bank"][3], torch.zeros_like(s0["bank"][3]))     # and the bank advanced


def _graphed_bank_worker(rank, world, port, out_dir):
    """The captured bank step on several ranks (graph_step.py: first half | ONE all_gather of the enqueue packet | second
    half) with the launch sequences run eagerly on the emulated kernels and the real gloo all_gather in between: replay r
    equals the eager trainer order (loss -> enqueue -> backward with the held-back write) on every rank, and the banks
    stay bit-identical across ranks."""
    import pytest
    import emu_harness
    import contrastiveseg_b200 as cs
    from contrastiveseg_b200 import bank as bank_mod, functional as Fn, graph_step
    from contrastiveseg_b200.synth import make_contrast_batch
    mp_ = pytest.MonkeyPatch()
    emu_harness.use_emulation(mp_)
    mp_.setattr(graph_step.GraphedContrastStep, "_capture", lambda self, warmup: None)
    mp_.setattr(graph_step.GraphedContrastStep, "_fork_zero_fill", lambda self: self._side_branch(0))
    mp_.setattr(graph_step.GraphedContrastStep, "_join_zero_fill", lambda self: None)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, D, M = 6, 64, 24
    data = make_contrast_batch(B=1, D=D, h=16, w=16, num_classes=K, img_stride=2, block=8, seed=300 + rank)
    embed, tgt, seg = data["embed"], data["target"], data["seg"]
    torch.manual_seed(0)
    bank_g, bank_e = cs.MemoryBank(K, M, D), cs.MemoryBank(K, M, D)
    bank_e.load_state_dict(bank_g.state_dict())
    opts = cs.ContrastOptions(temperature=0.07, base_temperature=0.07, max_samples=64, max_views=6, seed=5,
                              precision="fp32", num_classes=K)
    step = cs.GraphedContrastStep(embed, tgt, seg=seg, segment_queue=bank_g.segment_queue, pixel_queue=bank_g.pixel_queue,
                                  options=opts, enqueue=dict(bank=bank_g, network_stride=2, pixel_update_freq=4, seed=3))
    assert step.split and step.enq["world"] == world
    names = ("segment_queue", "pixel_queue", "segment_queue_ptr", "pixel_queue_ptr")
    for r in range(3):
        loss, grad = step.replay()
        loss, grad = loss.clone(), grad.clone()
        Fn._step_counter[0] = r
        bank_mod._enqueue_counter[0] = r
        e = embed.clone().requires_grad_(True)
        l = cs.pixel_contrast_loss(e, tgt, seg=seg, segment_queue=bank_e.segment_queue, pixel_queue=bank_e.pixel_queue,
                                   options=opts)
        bank_e.enqueue(e.detach(), tgt, network_stride=2, pixel_update_freq=4, seed=3)
        l.backward()
        assert torch.equal(l.detach(), loss) and torch.allclose(e.grad, grad, rtol=2e-6, atol=0)
        for n in names:
            assert torch.equal(getattr(bank_g, n), getattr(bank_e, n)), (r, n)
    torch.save([getattr(bank_g, n).clone() for n in names], os.path.join(out_dir, f"gb{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_graphed_bank_step_two_ranks_on_emulated_kernels(tmp_path):
    port = _free_port()
    mp.spawn(_graphed_bank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = torch.load(tmp_path / "gb0.pt"), torch.load(tmp_path / "gb1.pt")
    for x, y in zip(b0, b1):
        assert torch.equal(x, y)
    assert not torch.equal(b0[3], torch.zeros_like(b0[3]))
