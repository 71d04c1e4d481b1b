"""Import the UNMODIFIED reference modules for oracle validation / golden generation.

TEST INFRASTRUCTURE.  Works only where the reference tree is reachable (this build container:
/root/reference, $CSEG_REF, or a driver-provided baseline/_ref).  It never exists on the GPU box; nothing on the GPU path calls
this.  Recipe follows SURVEY.md appendix B.
"""
from __future__ import annotations

import ast
import os
import sys
import textwrap
import types
from typing import Optional

import torch
import torch.nn as nn


def reference_root() -> Optional[str]:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("CSEG_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "lib", "loss", "loss_contrast.py")):
            return cand
    return None


class DictConfiger:
    """Duck-typed Configer: the reference loss only calls get()/exists()."""

    def __init__(self, d):
        self.d = d

    def get(self, *keys):
        v = self.d
        for k in keys:
            v = v[k]
        return v

    def exists(self, *keys):
        v = self.d
        for k in keys:
            if not isinstance(v, dict) or k not in v:
                return False
            v = v[k]
        return True


_loaded = None


def load_reference():
    """Returns a namespace with .nomem, .mem (the loss modules) and .enqueue (the trainer's
    _dequeue_and_enqueue extracted by AST, since the trainer module itself is not importable
    here).  The reference hard-codes .cuda(); on a CPU-only host that is shimmed to identity."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not reachable")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import lib.loss.loss_contrast as nomem
        import lib.loss.loss_contrast_mem as mem
    src = open(os.path.join(root, "segmentor", "trainer_contrastive.py")).read()
    fn = next(n for c in ast.parse(src).body if isinstance(c, ast.ClassDef) and c.name == "Trainer"
              for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "_dequeue_and_enqueue")
    ns = {"torch": torch, "nn": nn}
    exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    _loaded = types.SimpleNamespace(nomem=nomem, mem=mem, enqueue=ns["_dequeue_and_enqueue"], root=root)
    return _loaded


class patched_randperm:
    """Context manager: route torch.randperm through perm_fn (recorder or replayer)."""

    def __init__(self, perm_fn):
        self.perm_fn = perm_fn

    def __enter__(self):
        self._orig = torch.randperm
        fn = self.perm_fn
        torch.randperm = lambda n, *a, **k: fn(int(n))
        return self

    def __exit__(self, *exc):
        torch.randperm = self._orig
        return False
