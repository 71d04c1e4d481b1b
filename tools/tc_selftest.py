"""GPU self-test of the tcgen05 pipeline (run on the B200 box): raw logit tiles against torch, then the fused
forward against the float64 oracle.  Each stage prints its error so one gpurun call localises a failure."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
torch.set_num_threads(min(16, __import__("os").cpu_count() or 1))

from contrastiveseg_b200 import functional as Fn
from contrastiveseg_b200.synth import make_sweep_point
from oracle import ref_port as P

dev = torch.device("cuda:0")
stage = sys.argv[1] if len(sys.argv) > 1 else "all"


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def dump_test(A, N):
    pt = make_sweep_point(A, N, D=256, num_classes=19, seed=A + N)
    a, c = pt["anchors"].to(dev), pt["contrast"].to(dev)
    c16 = Fn.to_bf16_rows(c, -(-N // 256) * 256)
    S = Fn.tc_dump_logits(a, c16, N)
    torch.cuda.synchronize()
    ref = bf(a).double() @ bf(c).double().t()
    err = (S.double() - ref).abs().max().item()
    print(f"dump A={A} N={N}: max|S-ref|={err:.3e}  (S[0,:4]={S[0,:4].tolist()} ref={ref[0,:4].tolist()})", flush=True)
    return err


def fwd_test(A, N, T=0.1, clustered=0.5, self_mode=False):
    pt = make_sweep_point(A, N, D=256, num_classes=19, seed=A * 7 + N, clustered=clustered)
    a, ya, c, yc = pt["anchors"], pt["ya"], pt["contrast"], pt["yc"]
    t0 = time.time()
    if self_mode:
        loss, st, _ = Fn.infonce_tc_forward(a.to(dev), ya.to(dev), temperature=T, base_temperature=0.07)
        cf = P.infonce_closed_form(bf(a).double(), ya, bf(a).double(), ya, T, 0.07, self_contrast=True)
        cf32 = P.infonce_closed_form(a.double(), ya, a.double(), ya, T, 0.07, self_contrast=True)
    else:
        c16 = Fn.to_bf16_rows(c.to(dev), -(-N // 256) * 256)
        diag = torch.arange(A)
        loss, st, _ = Fn.infonce_tc_forward(a.to(dev), ya.to(dev), contrast_bf16=c16, contrast_cls=yc.to(dev), n_cols=N,
                                            diag_col=diag.to(dev), temperature=T, base_temperature=0.07)
        cf = P.infonce_closed_form(bf(a).double(), ya, bf(c).double(), yc, T, 0.07, self_contrast=False)
        cf32 = P.infonce_closed_form(a.double(), ya, c.double(), yc, T, 0.07, self_contrast=False)
    torch.cuda.synchronize()
    l = loss.item()
    r16 = abs(l - cf["loss"].item()) / abs(cf["loss"].item())
    r32 = abs(l - cf32["loss"].item()) / abs(cf32["loss"].item())
    npos_ok = torch.equal(st[4].cpu().double(), cf["npos"])
    print(f"fwd A={A} N={N} self={self_mode}: loss={l:.6f} rel(bf16-oracle)={r16:.2e} rel(fp32-oracle)={r32:.2e} "
          f"npos_ok={npos_ok} ({time.time() - t0:.1f}s)", flush=True)
    return r16, r32


def bwd_test(A, N, T=0.1, clustered=0.5, self_mode=False):
    pt = make_sweep_point(A, N, D=256, num_classes=19, seed=A * 3 + N, clustered=clustered)
    a, ya, c, yc = pt["anchors"], pt["ya"], pt["contrast"], pt["yc"]
    if self_mode:
        loss, st, state = Fn.infonce_tc_forward(a.to(dev), ya.to(dev), temperature=T, base_temperature=0.07)
        cf = P.infonce_closed_form(bf(a).double(), ya, bf(a).double(), ya, T, 0.07, self_contrast=True)
    else:
        c16 = Fn.to_bf16_rows(c.to(dev), -(-N // 256) * 256)
        loss, st, state = Fn.infonce_tc_forward(a.to(dev), ya.to(dev), contrast_bf16=c16, contrast_cls=yc.to(dev), n_cols=N,
                                                diag_col=torch.arange(A).to(dev), temperature=T, base_temperature=0.07)
        cf = P.infonce_closed_form(bf(a).double(), ya, bf(c).double(), yc, T, 0.07, self_contrast=False)
    dA = Fn.infonce_tc_backward(state, st)
    torch.cuda.synchronize()
    ref = cf["dA"]
    err = (dA.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    fro = ((dA.cpu().double() - ref).norm() / ref.norm()).item()
    print(f"bwd A={A} N={N} self={self_mode}: max-abs err / max|g| = {err:.2e}  rel-Frobenius = {fro:.2e}  "
          f"(dA[0,:3]={dA[0,:3].tolist()} ref={ref[0,:3].tolist()})", flush=True)
    return err


if stage in ("all", "dump"):
    dump_test(128, 256)
    dump_test(200, 1000)
    dump_test(512, 4096)
if stage in ("all", "fwd"):
    fwd_test(200, 1000)
    fwd_test(912, 912, self_mode=True)
    fwd_test(1024, 20000, T=0.07)
    fwd_test(300, 190000 // 10, T=0.07, clustered=1.0)
if stage in ("all", "bwd"):
    bwd_test(128, 128 * 3)
    bwd_test(200, 1000)
    bwd_test(912, 912, self_mode=True)
    bwd_test(1024, 20000, T=0.07)
print("selftest done", flush=True)
