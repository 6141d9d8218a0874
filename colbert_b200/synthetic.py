"""Seeded synthetic ColBERT indexes, queries and candidate lists.

Used by the parity tests, the golden-vector generator (tests/golden/make_golden.py) and bench.py,
so that the reference (run in the authoring container), the oracle and the CUDA path all see
bit-identical inputs.  Everything is derived from ``numpy.random.default_rng(seed)`` (PCG64,
stream-stable) on the host; nothing here touches the GPU.

Layout written by :func:`write_index` is the reference's on-disk layout
(reference: colbert/indexing/encoder.py:139-149 writes, colbert/indexing/loaders.py:7-32 and
colbert/indexing/index_manager.py:12-18 read):  ``{part}.pt`` = ``torch.save`` of an fp16
``[N_part_tokens, dim]`` tensor and ``doclens.{part}.json`` = JSON list of ints.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class SynthIndex:
    """In-memory synthetic index: token embeddings (fp16, unit-norm rows) and per-doc lengths."""
    emb: np.ndarray          # [num_tokens, dim] float16
    doclens: np.ndarray      # [num_docs] int64
    part_sizes: List[int]    # number of docs in each part file

    @property
    def num_docs(self) -> int:
        return int(self.doclens.shape[0])

    @property
    def num_tokens(self) -> int:
        return int(self.emb.shape[0])

    @property
    def dim(self) -> int:
        return int(self.emb.shape[1])


def _unit_rows(x: np.ndarray) -> np.ndarray:
    n = np.linalg.norm(x, axis=1, keepdims=True)
    n[n == 0] = 1.0
    return x / n


def make_doclens(rng: np.random.Generator, num_docs: int, kind: str = "uniform",
                 lo: int = 1, hi: int = 180, fixed: Optional[int] = None) -> np.ndarray:
    """doclens per SURVEY.md §8d: 'uniform' U[lo,hi]; 'fixed' (multi-view: every doc has d_view rows)."""
    if kind == "fixed":
        assert fixed is not None
        return np.full(num_docs, int(fixed), dtype=np.int64)
    if kind == "uniform":
        return rng.integers(lo, hi + 1, size=num_docs, dtype=np.int64)
    if kind == "lognormal":
        v = np.exp(rng.normal(np.log(70.0), 0.5, size=num_docs))
        return np.clip(np.rint(v), lo, hi).astype(np.int64)
    raise ValueError(kind)


def make_index(seed: int, num_docs: int, dim: int = 128, doclen_kind: str = "uniform",
               lo: int = 1, hi: int = 180, fixed: Optional[int] = None, num_parts: int = 1,
               doclens: Optional[Sequence[int]] = None, normalize: bool = True,
               chunk_tokens: int = 1 << 20) -> SynthIndex:
    """randn → L2-normalise rows → fp16 (reference: BaseModel.py:26 normalises, encoder.py:175 halves)."""
    rng = np.random.default_rng(seed)
    if doclens is None:
        dl = make_doclens(rng, num_docs, doclen_kind, lo, hi, fixed)
    else:
        dl = np.asarray(doclens, dtype=np.int64)
        num_docs = dl.shape[0]
    total = int(dl.sum())
    emb = np.empty((total, dim), dtype=np.float16)
    for s in range(0, total, chunk_tokens):
        e = min(total, s + chunk_tokens)
        x = rng.standard_normal((e - s, dim), dtype=np.float32)
        if normalize:
            x = _unit_rows(x)
        emb[s:e] = x.astype(np.float16)
    # split docs into parts as evenly as possible
    base, rem = divmod(num_docs, num_parts)
    part_sizes = [base + (1 if i < rem else 0) for i in range(num_parts)]
    return SynthIndex(emb=emb, doclens=dl, part_sizes=part_sizes)


def make_queries(seed: int, num_queries: int, q_len: int = 32, dim: int = 128,
                 normalize: bool = True) -> np.ndarray:
    """Queries as fp32 [B, q_len, dim] with unit-norm rows."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((num_queries * q_len, dim), dtype=np.float32)
    if normalize:
        x = _unit_rows(x)
    return x.reshape(num_queries, q_len, dim).astype(np.float32)


def make_candidates(seed: int, num_queries: int, num_docs: int, per_query: int) -> np.ndarray:
    """[B, per_query] int64 distinct pids per query, uniform over the corpus, unsorted."""
    rng = np.random.default_rng(seed)
    out = np.empty((num_queries, per_query), dtype=np.int64)
    for b in range(num_queries):
        if per_query * 4 <= num_docs:
            # rejection-free draw of distinct ids for large corpora
            c = rng.choice(num_docs, size=per_query, replace=False, shuffle=True)
        else:
            c = rng.permutation(num_docs)[:per_query]
        out[b] = c
    return out


def write_index(index: SynthIndex, directory: str) -> None:
    """Write ``index`` in the reference's layout ({i}.pt + doclens.{i}.json)."""
    import torch
    os.makedirs(directory, exist_ok=True)
    doc0, tok0 = 0, 0
    for part, ndocs in enumerate(index.part_sizes):
        dl = index.doclens[doc0:doc0 + ndocs]
        ntok = int(dl.sum())
        t = torch.from_numpy(np.ascontiguousarray(index.emb[tok0:tok0 + ntok]))
        torch.save(t, os.path.join(directory, f"{part}.pt"))
        with open(os.path.join(directory, f"doclens.{part}.json"), "w") as f:
            json.dump([int(x) for x in dl], f)
        doc0 += ndocs
        tok0 += ntok
