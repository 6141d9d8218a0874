"""Import the reference's own ranker from oracle/_ref (staged by oracle/make_ref.py) — TEST / BASELINE
INFRASTRUCTURE ONLY; nothing under colbert_b200/ may import this.

The four import-time shims of SURVEY.md §8c, none of which touches a reference file:
  1. a stub ``faiss`` module (top-level import at colbert_ranker.py:4; only ColbertIndex uses it);
  2. ``ujson`` → the standard ``json`` (loaders.py:3);
  3. ``cpu=True``: colbert_ranker.DEVICE = 'cpu' (l.12) and Tensor.cuda → identity (l.112 hard-codes .cuda());
  4. ``cpu=True``: torch.zeros without pin_memory / device='cuda' (l.56-57).
With ``cpu=False`` (a GPU box) shims 3 and 4 are not applied and the reference runs the way its author deploys it:
CPU index_select into pinned staging buffers → H2D → einsum on the GPU.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "colbert", "ranking", "colbert_ranker.py"))


class _CpuShims:
    """Context manager for shims 3b and 4 (process-wide monkeypatches, undone on exit)."""

    def __enter__(self):
        import torch
        self.torch = torch
        self._zeros, self._cuda = torch.zeros, torch.Tensor.cuda

        def zeros(*a, **k):
            k.pop("pin_memory", None)
            if str(k.get("device", "cpu")).startswith("cuda"):
                k["device"] = "cpu"
            return self._zeros(*a, **k)

        torch.zeros = zeros
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        self.torch.zeros, self.torch.Tensor.cuda = self._zeros, self._cuda
        return False


def load(cpu: bool = True):
    """→ (colbert_ranker module, BaseModel class, shim context manager factory)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.modules.setdefault("faiss", types.ModuleType("faiss"))      # shim 1
    sys.modules.setdefault("ujson", json)                           # shim 2
    for name in [m for m in sys.modules if m == "colbert" or m.startswith("colbert.")]:
        del sys.modules[name]                                       # never mix with another `colbert` package
    cr = importlib.import_module("colbert.ranking.colbert_ranker")
    BaseModel = importlib.import_module("colbert.modeling.BaseModel").BaseModel
    cr.DEVICE = "cpu" if cpu else "cuda"                            # shim 3a
    return cr, BaseModel, (_CpuShims if cpu else _NoShims)


class _NoShims:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
