"""Timing of the fused seg-CE kernels vs the PyTorch ops at the Cityscapes shape (B=8, 19x128x256 -> 512x1024)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch, torch.nn.functional as F
import contrastiveseg_b200 as cs
from contrastiveseg_b200.synth import make_contrast_batch
dev = torch.device("cuda:0")
data = make_contrast_batch(B=8, D=8, h=128, w=256, num_classes=19, img_stride=4, block=32, seed=304)
seg, tgt = data["seg"].to(dev), data["target"].to(dev)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

s1 = seg.clone().requires_grad_(True)
def fused_fwd(): return cs.upsample_cross_entropy(s1, tgt, None, -1)
def fused_fb(): s1.grad = None; fused_fwd().backward()
def torch_fwd(): return F.cross_entropy(F.interpolate(s1, size=tgt.shape[1:], mode="bilinear", align_corners=True), tgt, ignore_index=-1)
def torch_fb(): s1.grad = None; torch_fwd().backward()
with torch.no_grad():
    print("fused fwd ms", timeit(fused_fwd), " torch fwd ms", timeit(torch_fwd))
print("fused fwd+bwd ms", timeit(fused_fb), " torch fwd+bwd ms", timeit(torch_fb))
