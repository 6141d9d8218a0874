"""Part-file (de)serialisation with the reference's names
(reference colbert/indexing/index_manager.py:4-18)."""
from __future__ import annotations

import torch


class IndexManager:
    """Writer side of the ``{i}.pt`` layout: ``save(tensor, path)`` is a plain ``torch.save``."""

    def __init__(self, dim=None):
        self.dim = dim

    def save(self, tensor: torch.Tensor, path_prefix: str) -> None:
        torch.save(tensor, path_prefix)


def load_index_part(filename: str, verbose: bool = True) -> torch.Tensor:
    """One part as a ``[N_i, dim]`` tensor on the host; legacy parts saved as a list of tensors are
    concatenated (reference index_manager.py:15-16)."""
    part = torch.load(filename, map_location="cpu")
    if isinstance(part, list):
        part = torch.cat(part)
    return part
