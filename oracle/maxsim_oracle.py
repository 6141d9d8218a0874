"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A numpy restatement of the late-interaction scoring path of wuyaoxuehun/colbert
(gather by doclen offset → MaxSim → top-k).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the product
(``colbert_b200``) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference holds exactly one known-answer vector for this path
(``BaseModel.test_score`` → ``[[21., 41.]]``, reference colbert/modeling/BaseModel.py:70-75); beyond
that, this oracle is pinned against outputs of the UNMODIFIED reference code executed in the
authoring container on seeded synthetic indexes (tests/golden/make_golden.py → tests/golden/*.npz,
checked by tests/test_oracle_golden.py).

The arithmetic of the reference path lives in a third-party dependency that is not vendored under
/root/reference: PyTorch (pinned ``torch==1.10.0``, reference requirements.txt:252 /
environment.yaml:58).  The call sites are ``torch.index_select`` (colbert_ranker.py:105),
``einsum/max/sum`` (BaseModel.py:43-45), ``sort`` (colbert_ranker.py:120,128) and ``kthvalue``
(colbert_ranker.py:241); their published semantics are restated below in numpy, in fp32.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import json
import os
from itertools import accumulate
from typing import List, Optional, Sequence, Tuple

import numpy as np

TAIL_PAD_ROWS = 512  # reference colbert_ranker.py:62 — zeros(num_embeddings + 512, dim)


# --------------------------------------------------------------------------------------------
# Store layout (reference rows a1-a4, a18)
# --------------------------------------------------------------------------------------------

def get_parts(directory: str):
    """reference colbert/indexing/loaders.py:7-19 — list ``*.pt``, sort AS INTEGERS, assert 0..P-1."""
    ext = ".pt"
    parts = sorted(int(fn[:-len(ext)]) for fn in os.listdir(directory) if fn.endswith(ext))
    assert list(range(len(parts))) == parts, parts
    parts_paths = [os.path.join(directory, f"{p}{ext}") for p in parts]
    samples_paths = [os.path.join(directory, f"{p}.sample") for p in parts]
    return parts, parts_paths, samples_paths


def load_doclens(directory: str, flatten: bool = True):
    """reference colbert/indexing/loaders.py:22-32 — one JSON int list per part."""
    parts, _, _ = get_parts(directory)
    all_doclens = []
    for p in parts:
        with open(os.path.join(directory, f"doclens.{p}.json")) as f:
            all_doclens.append(json.load(f))
    if flatten:
        all_doclens = [x for sub in all_doclens for x in sub]  # colbert/utils/utils.py:133-134
    return all_doclens


def load_store(directory: str, dim: int) -> Tuple[np.ndarray, np.ndarray]:
    """reference colbert_ranker.py:16-29,61-73 — one flat fp16 store of num_embeddings+512 rows
    (zero tail), parts copied in at running offsets.  Returns (store fp16, flat doclens int64)."""
    import torch  # only to parse the reference's torch.save container
    _, paths, _ = get_parts(directory)
    parts_doclens = load_doclens(directory, flatten=False)
    doclens = np.asarray([x for sub in parts_doclens for x in sub], dtype=np.int64)
    total = int(doclens.sum())
    store = np.zeros((total + TAIL_PAD_ROWS, dim), dtype=np.float16)
    off = 0
    for idx, path in enumerate(paths):
        end = off + int(sum(parts_doclens[idx]))
        part = torch.load(path)
        if isinstance(part, list):  # index_manager.py:15-16 legacy list-of-tensors
            part = torch.cat(part)
        store[off:end] = part.numpy()
        off = end
    return store, doclens


def pad_store(emb: np.ndarray) -> np.ndarray:
    """Append the 512 zero rows of colbert_ranker.py:62 to an in-memory embedding matrix."""
    out = np.zeros((emb.shape[0] + TAIL_PAD_ROWS, emb.shape[1]), dtype=emb.dtype)
    out[:emb.shape[0]] = emb
    return out


# --------------------------------------------------------------------------------------------
# init_ranker (reference rows a5-a7)
# --------------------------------------------------------------------------------------------

def doclens_pfxsum(doclens: Sequence[int]) -> np.ndarray:
    """reference colbert_ranker.py:32 — [0] + accumulate(doclens), int64 [N+1]."""
    return np.asarray([0] + list(accumulate(int(x) for x in doclens)), dtype=np.int64)


def percentile_kth(doclens: np.ndarray, p: int) -> int:
    """reference colbert_ranker.py:238-241 — ``kthvalue(int(p*len/100))``: the k-th SMALLEST value,
    k 1-indexed; k == 0 (len < 4 at p=25) raises in torch, here too."""
    assert p in range(1, 101)
    assert doclens.ndim == 1
    k = int(p * doclens.shape[0] / 100.0)
    if k < 1 or k > doclens.shape[0]:
        raise IndexError(f"kthvalue: k={k} out of range for {doclens.shape[0]} elements")
    return int(np.partition(doclens, k - 1)[k - 1])


def compute_strides(doclens: np.ndarray) -> List[int]:
    """reference colbert_ranker.py:36-40 — percentiles 25/50/75 ∪ {max}, de-duplicated, ascending."""
    doclens = np.asarray(doclens, dtype=np.int64)
    s = [percentile_kth(doclens, p) for p in (25, 50, 75)]
    s.append(int(doclens.max()))
    return sorted(set(s))


def bucket_assignments(doclens_sel: np.ndarray, strides: Sequence[int]) -> np.ndarray:
    """reference colbert_ranker.py:90 — bucket g = #{s in strides : doclen > s + 1e-6}."""
    st = np.asarray(strides, dtype=np.float64)[None, :]
    return (doclens_sel[:, None].astype(np.float64) > st + 1e-6).sum(-1)


def floor_flags(doclens_sel: np.ndarray, strides: Sequence[int]) -> np.ndarray:
    """SURVEY.md §8 a12′ — a doc sees padded (zeroed) slots in its stride bucket, and therefore a
    floor of 0 on every per-query-token max, iff its doclen is not itself one of the strides."""
    return ~np.isin(doclens_sel, np.asarray(strides, dtype=np.int64))


# --------------------------------------------------------------------------------------------
# BaseModel.score (reference row a12) — all-pairs MaxSim with multiplicative masks
# --------------------------------------------------------------------------------------------

def score_allpairs(Q: np.ndarray, D: np.ndarray, q_mask: np.ndarray, d_mask: np.ndarray) -> np.ndarray:
    """reference colbert/modeling/BaseModel.py:39-46:
        D = D * d_mask[..., None]; Q = Q * q_mask[..., None]
        simmat = einsum("qmh,dnh->qdmn"); max over n; sum over m   →  [q, d] fp32
    Masking is MULTIPLICATIVE: a masked slot contributes sim == 0 to the max, not -inf."""
    Q = Q.astype(np.float32) * q_mask[..., None].astype(np.float32)
    D = D.astype(np.float32) * d_mask[..., None].astype(np.float32)
    nq, m, h = Q.shape
    nd, n, _ = D.shape
    out = np.empty((nq, nd), dtype=np.float32)
    Dt = D.reshape(nd * n, h).T  # [h, d*n]
    for qi in range(nq):
        sim = (Q[qi] @ Dt).reshape(m, nd, n)          # [m, d, n]
        out[qi] = sim.max(-1).sum(0, dtype=np.float32)  # max over n, sum over m
    return out


def score_allpairs_grad(Q: np.ndarray, D: np.ndarray, q_mask: np.ndarray, d_mask: np.ndarray, W: np.ndarray):
    """What autograd derives for BaseModel.score (reference colbert/modeling/BaseModel.py:41-45) as the reference
    trains with it (colbert/modeling/colbert_model.py:87-95): for L = Σ_{q,d} W[q,d]·scores[q,d]

        n*(q,d,m) = argmax_n simmat[q,d,m,n]                    (first maximal n, like torch.max on the CPU)
        dL/dQ[q,m] = q_mask[q,m] · Σ_d W[q,d] · (D·d_mask)[d, n*(q,d,m)]
        dL/dD[d,n] = d_mask[d,n] · Σ_{(q,m): n*(q,d,m) = n} W[q,d] · (Q·q_mask)[q,m]

    → (scores [q,d], dQ [q,m,h], dD [d,n,h], argmax [q,d,m]), all float32 / int64."""
    qm = q_mask.astype(np.float32)
    dm = d_mask.astype(np.float32)
    Qm = Q.astype(np.float32) * qm[..., None]
    Dm = D.astype(np.float32) * dm[..., None]
    nq, m, h = Qm.shape
    nd, n, _ = Dm.shape
    scores = np.empty((nq, nd), dtype=np.float32)
    arg = np.empty((nq, nd, m), dtype=np.int64)
    dQ = np.zeros_like(Qm)
    dD = np.zeros_like(Dm)
    Dt = Dm.reshape(nd * n, h).T
    for qi in range(nq):
        sim = (Qm[qi] @ Dt).reshape(m, nd, n)
        a = sim.argmax(-1)                                   # [m, d]
        arg[qi] = a.T
        scores[qi] = np.take_along_axis(sim, a[..., None], -1)[..., 0].sum(0, dtype=np.float32)
        for di in range(nd):
            rows = Dm[di, a[:, di]]                          # [m, h]
            dQ[qi] += W[qi, di] * rows
            np.add.at(dD[di], a[:, di], W[qi, di] * Qm[qi])
    return scores, dQ * qm[..., None], dD * dm[..., None], arg


# --------------------------------------------------------------------------------------------
# rank_forward (reference rows a9-a15), faithful op sequence
# --------------------------------------------------------------------------------------------

def _stride_view(store: np.ndarray, stride: int) -> np.ndarray:
    """reference colbert_ranker.py:45-51 — as_strided(tensor, (T-stride+1, stride, dim), (dim, dim, 1)):
    view row o = store rows o..o+stride-1 (overlapping, zero-copy)."""
    T, dim = store.shape
    es = store.strides[1]
    return np.lib.stride_tricks.as_strided(store, shape=(T - stride + 1, stride, dim),
                                           strides=(dim * es, dim * es, es), writeable=False)


def rank_forward(store: np.ndarray, doclens: np.ndarray, pfxsum: np.ndarray, strides: Sequence[int],
                 Q: np.ndarray, pids: Sequence[int], depth: Optional[int] = 10,
                 output_D_embedding: bool = False, return_all_scores: bool = False):
    """reference colbert_ranker.py:75-137.

    ``Q`` is ``[1, dim, q_len]`` exactly as the reference receives it (faiss_indexers.py:232-234).
    Returns ``(pids, scores)`` as Python lists, score-descending, length ≤ depth; or, with
    ``output_D_embedding``, ``(pids, D fp32 [depth, stride, dim], mask bool [depth, stride])``.
    ``return_all_scores`` additionally returns the un-sorted fp32 scores in candidate order
    (the array the reference holds at colbert_ranker.py:122 before sorting)."""
    assert len(pids) > 0                                        # l.76
    assert Q.shape[0] in (1, len(pids))                         # l.77
    Qf = np.ascontiguousarray(Q).astype(np.float32)             # l.78
    raw_pids = [int(p) for p in pids]
    pids_a = np.asarray(raw_pids, dtype=np.int64)
    dl, offs = doclens[pids_a], pfxsum[pids_a]                  # l.88
    assign = bucket_assignments(dl, strides)                    # l.90

    out_scores, out_perm, out_D, out_mask = [], [], [], []
    one_to_n = np.arange(len(raw_pids))
    for g, stride in enumerate(strides):                        # l.96
        loc = assign == g
        if loc.sum() < 1e-5:
            continue
        g_dl, g_offs = dl[loc], offs[loc]
        # l.105-107: index_select on the stride-view (reads `stride` rows per doc: rows past the
        # doc's own length belong to the NEXT doc), then cast to fp32
        D = _stride_view(store, stride)[g_offs].astype(np.float32)
        mask = (np.arange(stride)[None, :] + 1) <= g_dl[:, None]    # l.108-109
        q = np.transpose(Qf, (0, 2, 1))                         # l.111 Q.permute(0,2,1) → [1, q_len, dim]
        s = score_allpairs(q, D, np.ones((1, q.shape[1]), dtype=np.int64), mask.astype(np.int64))[0]
        out_scores.append(s)
        out_perm.append(one_to_n[loc])
        out_D.append(D)
        out_mask.append(mask)

    perm = np.argsort(np.concatenate(out_perm), kind="stable")   # l.120
    scores = np.concatenate(out_scores)[perm]                    # l.122
    assert len(raw_pids) == scores.shape[0]                      # l.124-126

    # l.128-130: sort(descending=True) — torch's sort is unstable so the order of exact ties is
    # unspecified in the reference; the oracle (like the CUDA path) breaks ties by pid, lower first.
    order = np.lexsort((pids_a, -scores))
    top_pids = pids_a[order].tolist()[:depth]
    top_scores = scores[order].tolist()[:depth]
    if output_D_embedding:                                       # l.131-136 (single-bucket only)
        Dcat = np.concatenate(out_D)[perm]
        Mcat = np.concatenate(out_mask)[perm]
        res = (top_pids, Dcat[order][:depth], Mcat[order][:depth])
    else:
        res = (top_pids, top_scores)
    if return_all_scores:
        return res + (scores,)
    return res


# --------------------------------------------------------------------------------------------
# Independent formulation: exact-doclen MaxSim + the zero-floor rule (SURVEY.md §8 a12′)
# --------------------------------------------------------------------------------------------

def maxsim_exact(store: np.ndarray, doclens: np.ndarray, pfxsum: np.ndarray, strides: Sequence[int],
                 Qmd: np.ndarray, pids: Sequence[int], use_floor: bool = True) -> np.ndarray:
    """score_i = Σ_m max(max_{t<doclen_i} q_m·d_{i,t}, floor_i), floor_i = 0 if doclen_i ∉ strides
    else -inf.  ``Qmd`` is ``[q_len, dim]`` fp32.  This is what the CUDA kernel computes; it equals
    :func:`rank_forward`'s scores (checked in tests) without the stride-bucket over-read."""
    pids_a = np.asarray(pids, dtype=np.int64)
    dl, offs = doclens[pids_a], pfxsum[pids_a]
    fl = floor_flags(dl, strides) if use_floor else np.zeros(len(pids_a), dtype=bool)
    Qf = Qmd.astype(np.float32)
    out = np.empty(len(pids_a), dtype=np.float32)
    for i in range(len(pids_a)):
        D = store[offs[i]:offs[i] + dl[i]].astype(np.float32)   # [doclen, dim]
        mx = (Qf @ D.T).max(-1, initial=-np.inf)                 # [q_len]; an empty document has no rows: -inf
        if fl[i]:
            mx = np.maximum(mx, np.float32(0.0))
        out[i] = mx.sum(dtype=np.float32)
    return out


def topk_desc(scores: np.ndarray, ids: np.ndarray, k: Optional[int]):
    """Score-descending selection with the deterministic tie-break used by the CUDA path
    (equal scores: lower id first)."""
    order = np.lexsort((ids, -scores))
    if k is not None:
        order = order[:k]
    return ids[order], scores[order]


def gather_rows(store: np.ndarray, doclens: np.ndarray, pfxsum: np.ndarray, pids: Sequence[int],
                stride: int) -> Tuple[np.ndarray, np.ndarray]:
    """Rows a10/a15: ``stride`` consecutive store rows starting at each doc's offset (bit-exact
    copies, including the over-read into the next doc), cast to fp32, and the length mask."""
    pids_a = np.asarray(pids, dtype=np.int64)
    offs = pfxsum[pids_a]
    D = _stride_view(store, stride)[offs].astype(np.float32)
    mask = (np.arange(stride)[None, :] + 1) <= doclens[pids_a][:, None]
    return D, mask


def exhaustive_topk(store: np.ndarray, doclens: np.ndarray, pfxsum: np.ndarray, strides: Sequence[int],
                    Qmd: np.ndarray, k: int, use_floor: bool = True, pid_base: int = 0):
    """All documents scored for one query, then top-k."""
    n = doclens.shape[0]
    s = maxsim_exact(store, doclens, pfxsum, strides, Qmd, np.arange(n), use_floor=use_floor)
    ids, sc = topk_desc(s, np.arange(n, dtype=np.int64) + pid_base, k)
    return ids, sc


def merge_topk(list_scores: Sequence[np.ndarray], list_pids: Sequence[np.ndarray], k: int):
    """SURVEY.md §8e — merge W per-shard (score, pid) lists into the global top-k under the total
    order (score desc, pid asc); entries with pid < 0 are padding."""
    sc = np.concatenate(list_scores)
    pid = np.concatenate(list_pids)
    keep = pid >= 0
    sc, pid = sc[keep], pid[keep]
    order = np.lexsort((pid, -sc))[:k]
    return pid[order], sc[order]


def pack_keys(scores: np.ndarray, pids: np.ndarray) -> np.ndarray:
    """The 64-bit key the shards exchange (include/colbert_b200.h, cbk_topk_per_query_keys):
    [ordered(score) : 32 | ~pid : 32]; sorting keys descending = (score desc, pid asc); 0 = padding.
    -0.0 is canonicalised to +0.0 first."""
    s = (np.asarray(scores, dtype=np.float32) + np.float32(0.0)).view(np.uint32).astype(np.uint64)
    neg = (s & np.uint64(0x80000000)) != 0
    ordered = np.where(neg, (~s) & np.uint64(0xFFFFFFFF), s | np.uint64(0x80000000))
    low = (~np.asarray(pids, dtype=np.int64).astype(np.uint64)) & np.uint64(0xFFFFFFFF)
    return (ordered << np.uint64(32)) | low


def unpack_keys(keys: np.ndarray):
    """Inverse of :func:`pack_keys`; key 0 → (-inf, -1)."""
    keys = np.asarray(keys).astype(np.uint64)
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    pos = (hi & np.uint32(0x80000000)) != 0
    bits = np.where(pos, hi & np.uint32(0x7FFFFFFF), ~hi).astype(np.uint32)
    scores = bits.view(np.float32).copy()
    pids = ((~keys) & np.uint64(0xFFFFFFFF)).astype(np.int64)
    pad = keys == 0
    scores[pad] = -np.inf
    pids[pad] = -1
    return scores, pids
