"""Seeded input definitions shared by make_golden.py (which feeds them to the UNMODIFIED reference)
and by the tests (which feed the very same inputs to the oracle and to the CUDA path).
Only inputs live here; expected outputs live in the committed ``rank_*.npz`` fixtures."""
from __future__ import annotations

import numpy as np

from colbert_b200 import synthetic

CASES = [
    # plain rerank, variable doclens, several parts on disk
    dict(name="small", seed=11, num_docs=400, dim=128, doclen_kind="uniform", lo=1, hi=180, num_parts=3,
         q_lens=[32, 32, 32, 32], n_cand=150, depths=[("d10", 10), ("all", None)]),
    # very short docs + negative similarities: makes the reference's zero-floor (SURVEY §8 a12') visible,
    # and doclens that coincide with a stride (no floor for those)
    dict(name="edge", seed=23, num_docs=240, dim=128, doclen_kind="edge", num_parts=1,
         q_lens=[32, 32, 7], n_cand=240, depths=[("d10", 10), ("all", None)], negative_bias=True),
    # multi-view index: every doc has exactly d_view=8 embeddings, queries have q_view=8 rows
    dict(name="multiview", seed=37, num_docs=500, dim=128, doclen_kind="fixed", fixed=8, num_parts=2,
         q_lens=[8, 8, 8], n_cand=200, depths=[("d10", 10), ("all", None)], output_D=5),
    # exact ties: docs 2i and 2i+1 are identical
    dict(name="ties", seed=41, num_docs=120, dim=128, doclen_kind="uniform", lo=2, hi=40, num_parts=1,
         q_lens=[32, 16], n_cand=120, depths=[("all", None)], duplicate_pairs=True),
    # assorted query lengths, including > 32 rows
    dict(name="qlens", seed=53, num_docs=300, dim=128, doclen_kind="uniform", lo=1, hi=96, num_parts=1,
         q_lens=[1, 5, 17, 31, 48], n_cand=100, depths=[("all", None)]),
]


def _edge_doclens(rng: np.random.Generator, n: int) -> np.ndarray:
    dl = rng.integers(1, 61, size=n, dtype=np.int64)
    dl[: n // 4] = rng.integers(1, 4, size=n // 4)      # lots of 1..3-token docs
    dl[n // 4: n // 4 + 8] = 60                          # pin the max so it is a stride
    return rng.permutation(dl)


def build_case(case):
    """→ (SynthIndex, [Q_i fp32 [q_len_i, dim]], [pids_i int64 [n_cand]])"""
    seed = case["seed"]
    rng = np.random.default_rng(seed)
    kind = case["doclen_kind"]
    if kind == "edge":
        doclens = _edge_doclens(rng, case["num_docs"])
        index = synthetic.make_index(seed, case["num_docs"], case["dim"], doclens=doclens,
                                     num_parts=case["num_parts"])
    else:
        index = synthetic.make_index(seed, case["num_docs"], case["dim"], doclen_kind=kind,
                                     lo=case.get("lo", 1), hi=case.get("hi", 180), fixed=case.get("fixed"),
                                     num_parts=case["num_parts"])
    if case.get("duplicate_pairs"):
        # make doc 2i+1 a byte-for-byte copy of doc 2i
        dl = index.doclens.copy()
        dl[1::2] = dl[0::2]
        pf = np.concatenate([[0], np.cumsum(dl)])
        emb = np.empty((int(pf[-1]), index.dim), dtype=np.float16)
        src_pf = np.concatenate([[0], np.cumsum(index.doclens)])
        for i in range(0, case["num_docs"], 2):
            rows = index.emb[src_pf[i]: src_pf[i] + dl[i]]
            emb[pf[i]: pf[i + 1]] = rows
            emb[pf[i + 1]: pf[i + 2]] = rows
        index = synthetic.SynthIndex(emb=emb, doclens=dl, part_sizes=index.part_sizes)
    if case.get("negative_bias"):
        # push every embedding of the short docs into the negative orthant and the queries into
        # the positive one: all their similarities are < 0
        pf = np.concatenate([[0], np.cumsum(index.doclens)])
        emb = index.emb.astype(np.float32)
        for i in np.nonzero(index.doclens <= 3)[0]:
            x = -np.abs(emb[pf[i]: pf[i + 1]]) - 0.02
            emb[pf[i]: pf[i + 1]] = x / np.linalg.norm(x, axis=1, keepdims=True)
        index = synthetic.SynthIndex(emb=emb.astype(np.float16), doclens=index.doclens,
                                     part_sizes=index.part_sizes)
    queries, cands = [], []
    for qi, q_len in enumerate(case["q_lens"]):
        Q = synthetic.make_queries(seed * 1000 + qi, 1, q_len, case["dim"])[0]
        if case.get("negative_bias") and qi != 1:
            Q = np.abs(Q) + 0.02
            Q = (Q / np.linalg.norm(Q, axis=1, keepdims=True)).astype(np.float32)
        queries.append(Q)
        cands.append(synthetic.make_candidates(seed * 1000 + 500 + qi, 1, index.num_docs, case["n_cand"])[0])
    return index, queries, cands
