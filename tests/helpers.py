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
