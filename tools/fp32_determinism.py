"""Diagnostic: is the fp32 (SIMT) step bit-reproducible run to run, on fresh vs re-used workspaces?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contrastiveseg_b200 as cs
from contrastiveseg_b200 import functional as Fn
from contrastiveseg_b200.synth import make_contrast_batch

DEV = "cuda:0"
K, D = 7, 256
data = make_contrast_batch(B=2, D=D, h=32, w=32, num_classes=K, img_stride=4, block=16, seed=21)
embed, tgt, seg = data["embed"].to(DEV), data["target"].to(DEV), data["seg"].to(DEV)
for precision in ("fp32", "bf16"):
    opts = cs.ContrastOptions(temperature=0.07, base_temperature=0.07, max_samples=128, max_views=8, seed=5,
                              precision=precision, num_classes=K)
    outs = []
    for it in range(6):
        if it == 3:
            Fn.clear_workspaces()
        Fn._step_counter[0] = 0
        e = embed.clone().requires_grad_(True)
        l = cs.pixel_contrast_loss(e, tgt, seg=seg, options=opts)
        ws = Fn.last_workspace(e.device)
        l.backward()
        torch.cuda.synchronize()
        outs.append((l.detach().clone(), e.grad.clone(), ws.anchor_meta.clone(), ws.rowstats.clone(), ws.dA.clone(),
                     ws.anchors_f32.clone()))
    for it in range(1, 6):
        names = ("loss", "grad", "meta", "rowstats", "dA", "anchors")
        flags = []
        for n, a, b in zip(names, outs[0], outs[it]):
            eq = torch.equal(a, b)
            md = (a.float() - b.float()).abs().max().item() if not eq else 0.0
            flags.append(f"{n}={'==' if eq else '!= (%.3e)' % md}")
        print(precision, "run", it, " ".join(flags))
    hdr = Fn.last_workspace(torch.device(DEV)).plan_header()
    print(precision, "plan header", hdr)
