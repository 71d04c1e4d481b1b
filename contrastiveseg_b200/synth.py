"""Synthetic inputs S0..S4 of SURVEY.md §8(d) (no datasets / checkpoints in this environment).

All tensors are drawn from a seeded CPU ``torch.Generator`` so that every rank / arm / test sees
identical data for a given seed (the reference seeds 304, main_contrastive.py:154; per-rank seed is
304 + rank).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def block_label_map(B: int, Himg: int, Wimg: int, block: int, num_classes: int, g: torch.Generator,
                    ignore_label: int = -1) -> torch.Tensor:
    """(B,Himg,Wimg) int64 map of constant block x block squares, ids uniform in {ignore,0..K-1}."""
    hb, wb = -(-Himg // block), -(-Wimg // block)
    ids = torch.randint(-1, num_classes, (B, hb, wb), generator=g, dtype=torch.int64)
    ids = torch.where(ids < 0, torch.full_like(ids, ignore_label), ids)
    full = ids.repeat_interleave(block, dim=1).repeat_interleave(block, dim=2)
    return full[:, :Himg, :Wimg].contiguous()


def make_contrast_batch(B: int = 8, D: int = 256, h: int = 128, w: int = 256, num_classes: int = 19,
                        img_stride: int = 4, block: int = 32, boost: float = 2.0, seed: int = 304,
                        ignore_label: int = -1, himg: int = None, wimg: int = None):
    """S1-style batch.  Returns dict(embed_raw, embed, seg, target) — all CPU tensors.

    embed_raw ~ N(0,1) (B,D,h,w); embed = channel-L2-normalised; target = block label map at image
    resolution; seg = N(0,1) + boost * onehot(label at embedding resolution) so that roughly 57 %
    (boost 2.0) of labelled pixels are 'easy' (argmax == GT)."""
    g = _gen(seed)
    himg = himg if himg is not None else h * img_stride
    wimg = wimg if wimg is not None else w * img_stride
    target = block_label_map(B, himg, wimg, block, num_classes, g, ignore_label)
    embed_raw = torch.randn((B, D, h, w), generator=g, dtype=torch.float32)
    seg = torch.randn((B, num_classes, h, w), generator=g, dtype=torch.float32)
    lab = F.interpolate(target.unsqueeze(1).float(), size=(h, w), mode="nearest").squeeze(1).long()
    valid = (lab >= 0) & (lab < num_classes)
    onehot = F.one_hot(lab.clamp(0, num_classes - 1), num_classes).permute(0, 3, 1, 2).float()
    seg = seg + boost * onehot * valid.unsqueeze(1).float()
    embed = F.normalize(embed_raw, p=2, dim=1)
    return dict(embed_raw=embed_raw, embed=embed, seg=seg, target=target)


def make_bank(num_classes: int = 19, memory_size: int = 5000, dim: int = 256, seed: int = 305):
    """Bank buffers initialised like HRNet_W48_MEM.__init__ (lib/models/nets/hrnet.py:165-171)."""
    g = _gen(seed)
    seg_q = F.normalize(torch.randn((num_classes, memory_size, dim), generator=g), p=2, dim=2)
    pix_q = F.normalize(torch.randn((num_classes, memory_size, dim), generator=g), p=2, dim=2)
    return dict(segment_queue=seg_q, segment_queue_ptr=torch.zeros(num_classes, dtype=torch.long),
                pixel_queue=pix_q, pixel_queue_ptr=torch.zeros(num_classes, dtype=torch.long))


def make_sweep_point(A: int, N: int, D: int = 256, num_classes: int = 19, seed: int = 306,
                     clustered: float = 0.0, sorted_bank: bool = True):
    """S4: direct InfoNCE-level inputs.  anchors (A,D), bank rows (N,D) unit-norm; labels uniform over
    classes; bank labels sorted ascending (the bank is class-blocked in the real layout)."""
    g = _gen(seed)
    ya = torch.randint(0, num_classes, (A,), generator=g)
    yc = torch.randint(0, num_classes, (N,), generator=g)
    if sorted_bank:
        yc = torch.sort(yc).values
    centers = F.normalize(torch.randn((num_classes, D), generator=g), dim=1)
    a = torch.randn((A, D), generator=g) + clustered * centers[ya] * (D ** 0.5)
    c = torch.randn((N, D), generator=g) + clustered * centers[yc] * (D ** 0.5)
    return dict(anchors=F.normalize(a, dim=1), ya=ya, contrast=F.normalize(c, dim=1), yc=yc)
