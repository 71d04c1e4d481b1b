"""Host-side replay of the reference's sampling RNG (lib/loss/loss_contrast.py:79-82,
segmentor/trainer_contrastive.py:127): torch.randperm on the CPU generator, drawn in the reference's
data-dependent order.  Used by the 'torch_cpu' RNG mode (bit-identical samples to the reference for the
same torch seed, at the price of one small D2H copy per step) and by the parity tests (injected
permutations).  The default 'device' mode needs none of this."""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np
import torch


def host_plan(counts: np.ndarray, max_samples: int, max_views: int):
    """counts (B, 2K) int: [b, 2c] = hard pixels, [b, 2c+1] = easy pixels of class c in image b.
    Mirrors k_plan (csrc/pcl_select.cu) / loss_contrast.py:37-48,63-77.
    Returns (pairs [(b, c, n_hard, n_easy, keep_hard, keep_easy)], TC, V)."""
    B, NK = counts.shape
    K = NK // 2
    tot = counts[:, 0::2] + counts[:, 1::2]
    kept = [(b, c) for b in range(B) for c in range(K) if tot[b, c] > max_views]
    TC = len(kept)
    if TC == 0:
        return [], 0, 0
    V = min(max_samples // TC, max_views)
    pairs = []
    for b, c in kept:
        nh, ne = int(counts[b, 2 * c]), int(counts[b, 2 * c + 1])
        if 2 * nh >= V and 2 * ne >= V:
            kh = V // 2
            ke = V - kh
        elif 2 * nh >= V:
            ke = ne
            kh = V - ke
        elif 2 * ne >= V:
            kh = nh
            ke = V - kh
        else:
            raise RuntimeError(f"hard/easy split impossible: {nh} {ne} {V}")
        pairs.append((b, c, nh, ne, kh, ke))
    return pairs, TC, V


def anchor_rank_table(pairs, V: int, perm_fn: Callable[[int], torch.Tensor]) -> torch.Tensor:
    """(TC, V) int32: for every pair, the first keep_hard values of perm(n_hard) then the first keep_easy
    values of perm(n_easy) — both permutations are always drawn, in that order (loss_contrast.py:79-82)."""
    table = torch.zeros((max(len(pairs), 1), max(V, 1)), dtype=torch.int32)
    for t, (_, _, nh, ne, kh, ke) in enumerate(pairs):
        ph = perm_fn(nh)
        pe = perm_fn(ne)
        if V > 0:
            table[t, :kh] = ph[:kh].to(torch.int32)
            table[t, kh:kh + ke] = pe[:ke].to(torch.int32)
    return table


def bank_rank_table(counts: np.ndarray, F: int, perm_fn: Callable[[int], torch.Tensor]) -> torch.Tensor:
    """counts (B, K) pixels per (image, class) in the sub-sampled label grid.  One randperm(n) per
    (image asc, class asc, class > 0, n > 0) slot (trainer_contrastive.py:113-127); returns (B*K, F) int32."""
    B, K = counts.shape
    table = torch.zeros((B * K, max(F, 1)), dtype=torch.int32)
    for b in range(B):
        for c in range(1, K):
            n = int(counts[b, c])
            if n > 0:
                p = perm_fn(n)
                k = min(n, F)
                table[b * K + c, :k] = p[:k].to(torch.int32)
    return table
