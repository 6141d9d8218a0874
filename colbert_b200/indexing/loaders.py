"""Index-directory readers with the reference's names and return values
(reference colbert/indexing/loaders.py:7-32).

The on-disk contract: part files ``0.pt .. P-1.pt`` (``torch.save`` of an fp16 ``[N_i, dim]`` tensor),
one ``doclens.{i}.json`` (list of ints) per part; parts are ordered by their INTEGER value and must
be contiguous from 0.
"""
from __future__ import annotations

import json
import os
from typing import List, Tuple

PART_EXT = ".pt"


def get_parts(directory: str) -> Tuple[List[int], List[str], List[str]]:
    """→ (part numbers, ``{i}.pt`` paths, ``{i}.sample`` paths); asserts the numbering is 0..P-1."""
    numbers = []
    for entry in os.listdir(directory):
        if entry.endswith(PART_EXT):
            numbers.append(int(entry[: -len(PART_EXT)]))
    numbers.sort()                       # numeric, not lexicographic: 10.pt comes after 9.pt
    assert numbers == list(range(len(numbers))), numbers
    part_files = [os.path.join(directory, f"{i}{PART_EXT}") for i in numbers]
    sample_files = [os.path.join(directory, f"{i}.sample") for i in numbers]
    return numbers, part_files, sample_files


def load_doclens(directory: str, flatten: bool = True):
    """Per-part doclens lists (``flatten=False``) or one flat list in pid order (``flatten=True``)."""
    numbers, _, _ = get_parts(directory)
    per_part = []
    for i in numbers:
        with open(os.path.join(directory, f"doclens.{i}.json")) as fh:
            per_part.append(json.load(fh))
    if not flatten:
        return per_part
    flat: List[int] = []
    for lens in per_part:
        flat.extend(lens)
    return flat
