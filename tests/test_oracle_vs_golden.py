"""The CPU port (oracle) against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import ref_port as P
from helpers import load_golden, unpack_perms, rel_err

LOSS_CASES = ["nomem_small", "nomem_oddv", "nomem_mv1", "nomem_nondiv", "nomem_d256", "mem_small", "mem_d256"]


@pytest.mark.parametrize("name", LOSS_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_port_loss_and_grad_match_reference(name, dtype):
    g = load_golden(name)
    T, bT, ms, mv, K, ign = g["params"].tolist()
    embed = torch.from_numpy(g["embed"]).to(dtype).requires_grad_(True)
    queue = None
    if "segment_queue" in g:
        queue = torch.cat((torch.from_numpy(g["segment_queue"]), torch.from_numpy(g["pixel_queue"])), dim=1)
    replay = P.PermReplay(unpack_perms(g["perm_flat"], g["perm_lens"]))
    loss, det = P.pixel_contrast_loss(embed, torch.from_numpy(g["target"]), torch.from_numpy(g["predict"]),
                                      temperature=T, base_temperature=bT, max_samples=int(ms), max_views=int(mv),
                                      ignore_label=int(ign), queue=queue, perm_fn=replay, return_details=True)
    loss.backward()
    assert replay.pos == len(replay.draws)
    # sampled anchors are bit-identical (pure gather)
    TC, V = det["idx"].shape
    X_ = det["anchors"].detach().view(V, TC, -1).permute(1, 0, 2)
    if dtype == torch.float32:
        assert torch.equal(X_, torch.from_numpy(g["X_"]))
    assert torch.equal(det["cls"].float(), torch.from_numpy(g["y_"]))
    assert rel_err(loss.item(), g["loss"]) < 2e-6
    gref = torch.from_numpy(g["grad_embed"]).double()
    assert (embed.grad.double() - gref).abs().max().item() <= 2e-6 * gref.abs().max().item() + 1e-9


@pytest.mark.parametrize("name", ["nomem_small", "mem_small"])
def test_closed_form_matches_autograd(name):
    g = load_golden(name)
    T, bT, ms, mv, K, ign = g["params"].tolist()
    embed = torch.from_numpy(g["embed"]).double().requires_grad_(True)
    queue = None
    if "segment_queue" in g:
        queue = torch.cat((torch.from_numpy(g["segment_queue"]), torch.from_numpy(g["pixel_queue"])), dim=1).double()
    replay = P.PermReplay(unpack_perms(g["perm_flat"], g["perm_lens"]))
    loss, det = P.pixel_contrast_loss(embed, torch.from_numpy(g["target"]), torch.from_numpy(g["predict"]),
                                      temperature=T, base_temperature=bT, max_samples=int(ms), max_views=int(mv),
                                      queue=queue, perm_fn=replay, return_details=True)
    a = det["anchors"].detach().clone().requires_grad_(True)
    c = a if queue is None else det["contrast"]
    yc = det["ya"] if queue is None else det["yc"]
    l2 = P.infonce_dense(a, det["ya"], c, yc, T, bT)
    l2.backward()
    cf = P.infonce_closed_form(a.detach(), det["ya"], c.detach(), yc, T, bT, self_contrast=queue is None)
    assert rel_err(cf["loss"].item(), l2.item()) < 1e-12
    assert (cf["dA"] - a.grad).abs().max().item() < 1e-12


@pytest.mark.parametrize("name", ["enqueue_aligned", "enqueue_q6"])
def test_port_enqueue_matches_reference(name):
    g = load_golden(name)
    net_stride, M, Fq, steps, K = g["params"].tolist()
    sq, pq = torch.from_numpy(g["sq0"]).clone(), torch.from_numpy(g["pq0"]).clone()
    sp, pp = torch.zeros(K, dtype=torch.long), torch.zeros(K, dtype=torch.long)
    for s in range(steps):
        replay = P.PermReplay(unpack_perms(g[f"perm_flat{s}"], g[f"perm_lens{s}"]))
        P.dequeue_and_enqueue(torch.from_numpy(g[f"keys{s}"]), torch.from_numpy(g[f"labels{s}"]), sq, sp, pq, pp,
                              network_stride=net_stride, memory_size=M, pixel_update_freq=Fq, perm_fn=replay)
        assert torch.equal(sp, torch.from_numpy(g[f"sp{s + 1}"]))
        assert torch.equal(pp, torch.from_numpy(g[f"pp{s + 1}"]))
        assert torch.equal(pq, torch.from_numpy(g[f"pq{s + 1}"]))
        assert torch.allclose(sq, torch.from_numpy(g[f"sq{s + 1}"]), rtol=0, atol=1e-6)


@pytest.mark.parametrize("name,mem", [("wrapper_nomem_embed", False), ("wrapper_nomem_warmup", False),
                                      ("wrapper_mem_embed", True)])
def test_port_wrapper_matches_reference(name, mem):
    g = load_golden(name)
    T, bT, ms, mv, K, ign, lw, with_embed, _ = g["params"].tolist()
    seg = torch.from_numpy(g["seg"]).requires_grad_(True)
    embed = torch.from_numpy(g["embed"]).requires_grad_(True)
    preds = {"seg": seg, "embed": embed}
    if mem:
        preds["segment_queue"] = torch.from_numpy(g["segment_queue"])
        preds["pixel_queue"] = torch.from_numpy(g["pixel_queue"])
    replay = P.PermReplay(unpack_perms(g["perm_flat"], g["perm_lens"]))
    loss = P.contrast_ce_loss(preds, torch.from_numpy(g["target"]), with_embed=bool(with_embed), loss_weight=lw,
                              temperature=T, base_temperature=bT, max_samples=int(ms), max_views=int(mv),
                              ignore_label=int(ign), with_memory=mem, perm_fn=replay)
    loss.backward()
    assert rel_err(loss.item(), g["loss"]) < 2e-6
    assert torch.allclose(seg.grad, torch.from_numpy(g["grad_seg"]), rtol=1e-5, atol=1e-8)
    assert torch.allclose(embed.grad, torch.from_numpy(g["grad_embed"]), rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("name", ["wrapper_aux_embed", "wrapper_aux_warmup_weighted"])
def test_port_aux_wrapper_matches_reference(name):
    """ContrastAuxCELoss (lib/loss/loss_contrast.py:192-234) golden from the unmodified reference vs the port."""
    g = load_golden(name)
    T, bT, ms, mv, K, ign, lw, with_embed, w_seg, w_aux = g["params"].tolist()
    seg = torch.from_numpy(g["seg"]).requires_grad_(True)
    seg_aux = torch.from_numpy(g["seg_aux"]).requires_grad_(True)
    embed = torch.from_numpy(g["embed"]).requires_grad_(True)
    cw = torch.from_numpy(g["ce_weight"]) if g["ce_weight"].size else None
    replay = P.PermReplay(unpack_perms(g["perm_flat"], g["perm_lens"]))
    loss = P.contrast_auxce_loss({"seg": seg, "seg_aux": seg_aux, "embed": embed}, torch.from_numpy(g["target"]),
                                 with_embed=bool(with_embed), loss_weight=lw, seg_loss_weight=w_seg, aux_loss_weight=w_aux,
                                 temperature=T, base_temperature=bT, max_samples=int(ms), max_views=int(mv),
                                 ignore_label=int(ign), ce_weight=cw, perm_fn=replay)
    loss.backward()
    assert rel_err(loss.item(), g["loss"]) < 2e-6
    assert torch.allclose(seg.grad, torch.from_numpy(g["grad_seg"]), rtol=1e-5, atol=1e-8)
    assert torch.allclose(seg_aux.grad, torch.from_numpy(g["grad_seg_aux"]), rtol=1e-5, atol=1e-8)
    assert torch.allclose(embed.grad, torch.from_numpy(g["grad_embed"]), rtol=1e-4, atol=1e-8)
