"""GPU tests of features written after the round-1 GPU budget was spent (the file sorts last among the GPU tests on
purpose: it contains the CUDA-graph capture).  They pass on the CPU emulator (tests/test_emu_kernels.py) and the host
logic is covered by tests/test_host_wiring.py; what is left to confirm on hardware is listed per test.  Once they have
passed on a B200 they move into test_gpu_parity.py.  See `pytestmark` for how they are reported until then."""
import os

import pytest
import torch

import contrastiveseg_b200 as cs
from contrastiveseg_b200 import functional as Fn
from contrastiveseg_b200.synth import make_bank, make_contrast_batch

# Not yet run on hardware when committed (the kernels and the host logic pass on the CPU emulator of tests/emu).  Until
# they have passed once on a B200 (then: PCL_TEST_EXPERIMENTAL=1 makes them strict) a failure is reported as "xfailed"
# and a success as "xpassed" — the verified `-m gpu` suite stays a clean signal either way; the timeout bounds a surprise.
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300),
              pytest.mark.xfail(condition=not os.environ.get("PCL_TEST_EXPERIMENTAL"), strict=False,
                                reason="first hardware run pending (passes on the CPU emulator, tests/test_emu_kernels.py)")]
DEV = "cuda:0"


def _cfg(K, precision="fp32", **extra):
    c = {"temperature": 0.07, "base_temperature": 0.07, "max_samples": 128, "max_views": 8, "loss_weight": 0.1,
         "precision": precision}
    c.update(extra)
    return cs.Configer({"data": {"num_classes": K}, "contrast": c, "loss": {"params": {"ce_ignore_index": -1}},
                        "network": {"stride": 4}})


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_enqueue_between_loss_and_backward_does_not_change_the_gradient(precision):
    """Trainer order loss -> enqueue -> backward (trainer_contrastive.py:241-255): the reference's autograd holds a
    copy of the bank, the engine re-reads it in the backward sweep, so the in-place write is held back until after
    that backward.  Property: the gradient is bit-identical to the run without an enqueue, and the bank afterwards is
    bit-identical to an immediate enqueue."""
    K, D, M = 6, 256, 32
    data = make_contrast_batch(B=2, D=D, h=16, w=32, num_classes=K, img_stride=4, block=16, seed=5)
    tgt, seg = data["target"].to(DEV), data["seg"].to(DEV)
    banks = []
    grads = []
    for with_enqueue in (False, True):
        torch.manual_seed(0)
        bank = cs.MemoryBank(K, M, D, with_shadow=(precision == "bf16")).to(DEV)
        if precision == "bf16":
            bank.sync_shadow()
        crit = cs.PixelContrastLoss(_cfg(K, precision, seed=11))
        Fn._step_counter[0] = 0                                   # same device-RNG stream in both runs
        embed = data["embed"].to(DEV).requires_grad_(True)
        loss = crit(embed, tgt, seg=seg, queue=(bank.segment_queue, bank.pixel_queue), bank_shadow=bank.shadow)
        before = bank.pixel_queue.clone()
        if with_enqueue:
            bank.enqueue(embed.detach(), tgt, network_stride=4, pixel_update_freq=5, seed=3)
            assert torch.equal(bank.pixel_queue, before)          # held back: a backward that reads the bank is pending
        loss.backward()
        if with_enqueue:
            assert not torch.equal(bank.pixel_queue, before)      # ... and written right after it
        grads.append(embed.grad.clone())
        banks.append(bank)
    assert torch.equal(grads[0], grads[1])
    # the deferred write equals an immediate one on a fresh copy of the initial bank
    torch.manual_seed(0)
    ref = cs.MemoryBank(K, M, D, with_shadow=(precision == "bf16")).to(DEV)
    if precision == "bf16":
        ref.sync_shadow()
    from contrastiveseg_b200 import bank as bank_mod
    bank_mod._enqueue_counter[0] -= 1                             # replay the same enqueue seed
    ref.enqueue(data["embed"].to(DEV), tgt, network_stride=4, pixel_update_freq=5, seed=3)
    for name in ("segment_queue", "pixel_queue", "segment_queue_ptr", "pixel_queue_ptr"):
        assert torch.equal(getattr(ref, name), getattr(banks[1], name)), name
    if precision == "bf16":
        assert torch.equal(ref.shadow, banks[1].shadow)


@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("precision,mem", [("bf16", False), ("fp32", False), ("bf16", True)])
def test_graphed_step_equals_eager_step(precision, mem, overlap):
    """GraphedContrastStep: replay r samples like the eager step with counter r+1 (same keyed bijection), so anchors,
    loss and dense gradient are bit-identical; successive replays draw different anchors; grad_scale scales the gradient.
    To confirm on hardware: capture of the whole C-ABI sequence (no illegal call under capture)."""
    K, D, M = 7, 256, 48
    data = make_contrast_batch(B=2, D=D, h=32, w=32, num_classes=K, img_stride=4, block=16, seed=21)
    embed, tgt, seg = data["embed"].to(DEV), data["target"].to(DEV), data["seg"].to(DEV)
    bank = None
    kw = {}
    if mem:
        bank = cs.MemoryBank(K, M, D, with_shadow=True).to(DEV)
        bank.sync_shadow()
        kw = dict(segment_queue=bank.segment_queue, pixel_queue=bank.pixel_queue, bank_shadow=bank.shadow)
    opts = cs.ContrastOptions(temperature=0.07, base_temperature=0.07, max_samples=128, max_views=8, seed=5,
                              precision=precision, num_classes=K)
    step = cs.GraphedContrastStep(embed, tgt, seg=seg, options=opts, overlap_zero_fill=overlap, **kw)
    metas, losses, grads = [], [], []
    for r in range(3):
        loss, grad = step.replay()
        torch.cuda.synchronize()
        metas.append(step.ws.anchor_meta.clone()); losses.append(loss.clone()); grads.append(grad.clone())
    assert int(step.counter.item()) == 3
    assert not torch.equal(metas[0], metas[1]) and not torch.equal(metas[1], metas[2])
    for r in range(3):
        Fn._step_counter[0] = r                                   # the eager call pre-increments: counter r+1
        e = embed.clone().requires_grad_(True)
        l = cs.pixel_contrast_loss(e, tgt, seg=seg, options=opts, **{k: v for k, v in kw.items()})
        ws = Fn.last_workspace(e.device)
        l.backward()
        torch.cuda.synchronize()
        assert torch.equal(ws.anchor_meta, metas[r])
        assert torch.equal(l.detach(), losses[r])
        if overlap:    # rows reduced by the sweep's own reduction kernel instead of the fused writer: same sum order
            assert torch.allclose(e.grad, grads[r], rtol=2e-6, atol=0)
        else:
            assert torch.equal(e.grad, grads[r])
    step.set_grad_scale(0.25)
    step.counter.fill_(1)
    _, g = step.replay()
    torch.cuda.synchronize()
    assert torch.allclose(g, 0.25 * grads[1], rtol=1e-6, atol=0)
    # autograd hand-off
    e = embed.clone().requires_grad_(True)
    step2 = cs.GraphedContrastStep(e, tgt, seg=seg, options=opts, **kw)
    out = step2.apply(e)
    out.backward()
    assert torch.equal(e.grad, step2.grad)
