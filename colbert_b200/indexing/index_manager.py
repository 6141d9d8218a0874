"""Part files of the index: one ``{i}.pt`` per part holding the per-token embeddings ``[N_i, dim]``
(names and call signatures of reference colbert/indexing/index_manager.py:4-18).

The loader is what feeds the HBM-resident store, so it checks what the kernels later rely on (2-D, floating point,
a single width per index) instead of failing inside a launch, and can memory-map a part so that a 100 GB index is
streamed into the device buffer without a second host copy.
"""
from __future__ import annotations

import os
from typing import Optional

import torch


def _as_part_tensor(obj, filename: str) -> torch.Tensor:
    # parts written by old encoders are Python lists of per-batch tensors (the reference concatenates them too)
    if isinstance(obj, (list, tuple)):
        if len(obj) == 0:
            raise ValueError(f"{filename}: empty part")
        obj = torch.cat([torch.as_tensor(x) for x in obj])
    if not isinstance(obj, torch.Tensor):
        raise TypeError(f"{filename}: expected a tensor (or a list of tensors), found {type(obj).__name__}")
    if obj.dim() != 2 or not obj.is_floating_point():
        raise ValueError(f"{filename}: expected a floating-point [rows, dim] matrix, found {tuple(obj.shape)} {obj.dtype}")
    return obj


ALLOW_PICKLE_ENV = "COLBERT_B200_ALLOW_PICKLE"


def load_index_part(filename: str, verbose: bool = True, mmap: bool = False, dim: Optional[int] = None,
                    allow_pickle: Optional[bool] = None) -> torch.Tensor:
    """The embeddings of one part as a host tensor ``[N_i, dim]``.  ``mmap``: map the file instead of reading it
    (zip-format checkpoints only); ``dim``: width the caller expects, checked here.

    Parts are read with the weights-only unpickler (tensors and lists of tensors — both forms the reference writes —
    load under it).  A file it refuses is NOT retried with the full pickle machinery unless the caller opts in with
    ``allow_pickle=True`` (or ``COLBERT_B200_ALLOW_PICKLE=1`` in the environment): a part file is data, and a crafted
    one must not get code execution just by failing the safe load."""
    kwargs = {"map_location": "cpu"}
    if mmap:
        kwargs["mmap"] = True
    if allow_pickle is None:
        allow_pickle = os.environ.get(ALLOW_PICKLE_ENV, "0") == "1"
    try:
        obj = torch.load(filename, weights_only=True, **kwargs)
    except Exception as exc:
        if not allow_pickle:
            raise RuntimeError(
                f"{filename}: the weights-only loader refused this part ({type(exc).__name__}: {exc}). If it is a "
                f"trusted legacy pickle, pass allow_pickle=True or set {ALLOW_PICKLE_ENV}=1.") from exc
        obj = torch.load(filename, weights_only=False, **kwargs)
    part = _as_part_tensor(obj, filename)
    if dim is not None and part.size(1) != dim:
        raise ValueError(f"{filename}: embeddings are {part.size(1)} wide, the ranker was built for dim={dim}")
    if verbose:
        print(f"#> loaded {os.path.basename(filename)}: {part.size(0)} embeddings x {part.size(1)} ({part.dtype})", flush=True)
    return part


class IndexManager:
    """Writer side of the same layout (the reference's encoder calls ``IndexManager(dim).save(embs, path)``)."""

    def __init__(self, dim=None):
        self.dim = dim

    def save(self, tensor: torch.Tensor, path_prefix: str) -> None:
        part = _as_part_tensor(tensor, path_prefix)
        if self.dim is not None and part.size(1) != self.dim:
            raise ValueError(f"refusing to write a {part.size(1)}-wide part into a dim={self.dim} index")
        torch.save(part.contiguous(), path_prefix)
