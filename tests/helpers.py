"""Shared helpers for the parity tests (golden loading, permutation replay)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def unpack_perms(flat, lens):
    out, at = [], 0
    for n in lens.tolist():
        out.append(torch.from_numpy(flat[at:at + n].astype(np.int64)))
        at += n
    return out


def rel_err(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-30)


def bank_infonce_chunked(anchors, ya, diag, segq, pixq, T, bT, chunk=32768):
    """Column-chunked float64 closed form of the bank-mode InfoNCE (lib/loss/loss_contrast_mem.py:
    91-152) in plain torch ops, for contrast sets too large for a dense A x N matrix (configs[2]: N = 190 000,
    configs[3]: N = 1.71 M).  Contrast rows: classes 1..K-1 as blocks of R = M0 + M1 rows (segment rows, then pixel
    rows), then R all-zero rows of label 0 (Q2/Q3); the (i, diag_i) entry is removed from the positives (Q1).
    anchors (A,D), ya (A,) int, diag (A,) int (reference row index), segq/pixq (K,M,D).  Runs on anchors.device.
    Returns dict(loss, dA, m, neg, npos).  Checked against oracle.ref_port.infonce_closed_form on small sizes
    (tests/test_host_logic.py)."""
    dev = anchors.device
    a = anchors.double()
    A, D = a.shape
    K, M0 = segq.shape[0], segq.shape[1]
    M1 = pixq.shape[1]
    R = M0 + M1
    n_real = (K - 1) * R
    N = n_real + R
    ya = ya.to(dev).long()
    diag = diag.to(dev).long()

    def block(c0, c1):
        """rows [c0, c1) of the flattened contrast matrix and their labels"""
        cols = torch.arange(c0, c1, device=dev)
        lab = torch.where(cols < n_real, cols // R + 1, torch.zeros_like(cols))
        out = torch.zeros((c1 - c0, D), dtype=torch.float64, device=dev)
        real = cols < n_real
        cr = cols[real]
        cls = cr // R + 1
        q = cr - (cls - 1) * R
        seg_rows = q < M0
        rows = torch.empty((cr.numel(), D), dtype=torch.float64, device=dev)
        rows[seg_rows] = segq[cls[seg_rows], q[seg_rows]].double()
        rows[~seg_rows] = pixq[cls[~seg_rows], q[~seg_rows] - M0].double()
        out[real] = rows
        return out, lab, cols

    m = torch.full((A,), -float("inf"), dtype=torch.float64, device=dev)
    for c0 in range(0, N, chunk):
        C, lab, cols = block(c0, min(N, c0 + chunk))
        m = torch.maximum(m, ((a @ C.t()) / T).max(1).values)
    neg = torch.zeros(A, dtype=torch.float64, device=dev)
    for c0 in range(0, N, chunk):
        C, lab, cols = block(c0, min(N, c0 + chunk))
        e = torch.exp((a @ C.t()) / T - m[:, None])
        neg += (e * (lab[None, :] != ya[:, None])).sum(1)
    possum = torch.zeros(A, dtype=torch.float64, device=dev)
    s = torch.zeros_like(possum)
    npos = torch.zeros_like(possum)
    for c0 in range(0, N, chunk):
        C, lab, cols = block(c0, min(N, c0 + chunk))
        lm = (a @ C.t()) / T - m[:, None]
        e = torch.exp(lm)
        pos = (lab[None, :] == ya[:, None]) & (cols[None, :] != diag[:, None])
        possum += ((lm - torch.log(e + neg[:, None])) * pos).sum(1)
        s += (pos / (e + neg[:, None])).sum(1)
        npos += pos.sum(1)
    loss = (-(T / bT) * possum / npos).mean()
    c = (T / bT) / (A * npos)
    dA = torch.zeros((A, D), dtype=torch.float64, device=dev)
    for c0 in range(0, N, chunk):
        C, lab, cols = block(c0, min(N, c0 + chunk))
        e = torch.exp((a @ C.t()) / T - m[:, None])
        same = lab[None, :] == ya[:, None]
        pos = same & (cols[None, :] != diag[:, None])
        inv = 1.0 / (e + neg[:, None])
        G = torch.where(pos, -c[:, None] * (1 - e * inv), torch.zeros_like(e)) + \
            torch.where(~same, c[:, None] * e * s[:, None], torch.zeros_like(e))
        dA += (G @ C) / T
    return dict(loss=loss, dA=dA, m=m, neg=neg, npos=npos)
