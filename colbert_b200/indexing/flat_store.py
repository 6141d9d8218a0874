"""Flat binary form of the index store for fast (sharded) loading into HBM — SURVEY.md §8f #1.

The reference keeps an index as ``{i}.pt`` (``torch.save`` of fp16 ``[N_i, dim]``) + ``doclens.{i}.json``
(written colbert/indexing/encoder.py:139-149, read colbert/indexing/loaders.py:7-32 and
colbert/indexing/index_manager.py:12-18); loading 160 GB through ``torch.load`` + JSON takes minutes and
cannot read a pid range without parsing every part.  ``convert_index`` rewrites it ONCE as

    store.bin      raw row-major 16-bit matrix [num_embeddings, dim]          (mmap-able; no header)
    doclens.i32    raw little-endian int32 [num_docs]
    meta.json      {"dim", "dtype", "num_docs", "num_embeddings", "parts"}

and ``load_flat`` maps any contiguous pid range (a shard, colbert_b200/sharding.py) straight into a device
tensor in the reference's in-memory layout (num rows + 512 zero tail rows, colbert_ranker.py:62).
"""
from __future__ import annotations

import json
import os
from typing import Optional, Tuple

import numpy as np
import torch

from .index_manager import load_index_part
from .loaders import get_parts, load_doclens

_DTYPES = {"float16": (np.float16, torch.float16), "bfloat16": (np.uint16, torch.bfloat16)}


def convert_index(index_path: str, out_path: str) -> dict:
    """Reference layout → flat layout (streams one part at a time)."""
    os.makedirs(out_path, exist_ok=True)
    _, part_files, _ = get_parts(index_path)
    parts_doclens = load_doclens(index_path, flatten=False)
    doclens = np.asarray([x for sub in parts_doclens for x in sub], dtype=np.int32)
    dim, dtype_name, rows = None, None, 0
    with open(os.path.join(out_path, "store.bin"), "wb") as out:
        for path, dl in zip(part_files, parts_doclens):
            part = load_index_part(path, verbose=False)
            assert part.size(0) == sum(dl), (path, part.size(0), sum(dl))
            name = str(part.dtype).replace("torch.", "")
            assert name in _DTYPES, f"unsupported part dtype {part.dtype}"
            dim = dim or part.size(1)
            dtype_name = dtype_name or name
            assert part.size(1) == dim and name == dtype_name
            out.write(part.contiguous().view(torch.int16).numpy().tobytes())
            rows += part.size(0)
    doclens.tofile(os.path.join(out_path, "doclens.i32"))
    meta = {"dim": int(dim), "dtype": dtype_name, "num_docs": int(doclens.shape[0]), "num_embeddings": int(rows),
            "parts": len(part_files)}
    with open(os.path.join(out_path, "meta.json"), "w") as fh:
        json.dump(meta, fh)
    return meta


def read_meta(flat_path: str) -> Tuple[dict, np.ndarray]:
    with open(os.path.join(flat_path, "meta.json")) as fh:
        meta = json.load(fh)
    doclens = np.fromfile(os.path.join(flat_path, "doclens.i32"), dtype=np.int32)
    assert doclens.shape[0] == meta["num_docs"] and int(doclens.sum()) == meta["num_embeddings"]
    return meta, doclens


def load_flat(flat_path: str, device, pid_lo: int = 0, pid_hi: Optional[int] = None, chunk_rows: int = 1 << 22):
    """→ (store [rows + 512, dim] on ``device``, doclens int64 [pid_hi - pid_lo], pid_lo, meta).  Only the bytes of
    the requested pid range are read (memory-mapped, copied through a pinned bounce buffer)."""
    meta, doclens = read_meta(flat_path)
    pid_hi = meta["num_docs"] if pid_hi is None else pid_hi
    pfx = np.concatenate([[0], np.cumsum(doclens, dtype=np.int64)])
    r0, r1 = int(pfx[pid_lo]), int(pfx[pid_hi])
    np_dt, t_dt = _DTYPES[meta["dtype"]]
    dim = meta["dim"]
    mm = np.memmap(os.path.join(flat_path, "store.bin"), dtype=np.uint16, mode="r", shape=(meta["num_embeddings"], dim))
    store = torch.zeros((r1 - r0 + 512, dim), dtype=t_dt, device=device)
    bounce = torch.empty((min(chunk_rows, max(1, r1 - r0)), dim), dtype=torch.int16).pin_memory() \
        if torch.device(device).type == "cuda" else None
    for s in range(r0, r1, chunk_rows):
        e = min(r1, s + chunk_rows)
        src = torch.from_numpy(np.ascontiguousarray(mm[s:e]).view(np.int16))
        dst = store[s - r0: e - r0].view(torch.int16)
        if bounce is not None:
            bounce[: e - s].copy_(src)
            dst.copy_(bounce[: e - s], non_blocking=False)
        else:
            dst.copy_(src)
    return store, torch.from_numpy(doclens[pid_lo:pid_hi].astype(np.int64)), pid_lo, meta
