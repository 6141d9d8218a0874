"""Randomised parity sweep (seeded): many small, oddly-shaped problems through every scoring kernel
(mma.sync rerank, tcgen05 rerank, tcgen05 exhaustive) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as O
from parity_utils import SCORE_RTOL

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = torch.device("cuda", 0)


def _random_problem(rng):
    from colbert_b200 import synthetic
    n_docs = int(rng.integers(1, 400))
    kind = rng.integers(0, 4)
    if kind == 0:
        doclens = rng.integers(1, 181, size=n_docs)
    elif kind == 1:
        doclens = np.full(n_docs, int(rng.integers(1, 20)))              # multi-view-like: all equal
    elif kind == 2:
        doclens = rng.integers(1, 4, size=n_docs)                         # tiny documents
    else:
        doclens = rng.integers(100, 400, size=n_docs)                     # long documents (several tiles)
    index = synthetic.make_index(int(rng.integers(1 << 30)), n_docs, dim=128, doclens=doclens.astype(np.int64))
    n_q = int(rng.integers(1, 9))
    q_len = int(rng.integers(1, 33))
    Q = synthetic.make_queries(int(rng.integers(1 << 30)), n_q, q_len, 128)
    lens = rng.integers(0, 60, size=n_q)
    cands = [rng.integers(0, n_docs, size=l).astype(np.int64) for l in lens]
    return index, Q, cands


@pytest.mark.parametrize("seed", range(12))
def test_random_problems_all_kernels(seed):
    from colbert_b200 import _lib
    from colbert_b200.ranking import ColbertRanker
    rng = np.random.default_rng(1000 + seed)
    index, Q, cands = _random_problem(rng)
    dt = torch.bfloat16 if seed % 2 else torch.float16
    emb = torch.from_numpy(index.emb).to(dt)
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
    store, pf = O.pad_store(emb.float().numpy()), O.doclens_pfxsum(index.doclens)
    flat = np.concatenate(cands) if sum(len(c) for c in cands) else np.zeros(0, np.int64)
    rowptr = np.concatenate([[0], np.cumsum([len(c) for c in cands])]).astype(np.int64)
    refs = [O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], cands[b]) if len(cands[b]) else np.zeros(0, np.float32)
            for b in range(len(cands))]
    ref_flat = np.concatenate(refs) if flat.size else np.zeros(0, np.float32)
    Qd = torch.from_numpy(Q).to(DEV)
    if flat.size:
        for flags in (0, _lib.CBK_FLAG_RERANK_TCGEN05):
            ranker.kernel_flags = flags
            got = ranker.score_candidates(Qd, torch.from_numpy(flat).to(DEV), torch.from_numpy(rowptr).to(DEV)).cpu().numpy()
            rel = np.abs(got - ref_flat) / np.maximum(np.abs(ref_flat), 1.0)
            assert rel.max() <= SCORE_RTOL, (seed, flags, rel.max())
    ranker.kernel_flags = 0
    dense = ranker.score_all(Qd).cpu().numpy()
    full = np.stack([O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], np.arange(index.num_docs)) for b in range(Q.shape[0])])
    rel = np.abs(dense - full) / np.maximum(np.abs(full), 1.0)
    assert rel.max() <= SCORE_RTOL, (seed, "exhaustive", rel.max())


@pytest.mark.parametrize("dim", [128, 64])
def test_empty_documents_score_zero(dim):
    """Documents of length 0 among the candidates (the reference scores them 0, see tests/test_oracle_properties.py):
    every rerank kernel — mma.sync, tcgen05, generic width — and the full rank_forward call."""
    from colbert_b200 import _lib, synthetic
    from colbert_b200.ranking import ColbertRanker
    index = synthetic.make_index(977 + dim, 400, dim=dim, lo=0, hi=14)
    n_empty = int((index.doclens == 0).sum())
    assert 0 < n_empty < 90
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=DEV)
    assert 0 not in ranker.strides
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    Q = synthetic.make_queries(978, 2, 32, dim)
    pids = np.arange(400, dtype=np.int64)
    flat = np.concatenate([pids, pids[::-1]])
    rowptr = np.array([0, 400, 800], dtype=np.int64)
    ref = np.concatenate([O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[0], pids),
                          O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[1], pids[::-1])])
    for flags in ((0, _lib.CBK_FLAG_RERANK_TCGEN05) if dim == 128 else (0,)):
        ranker.kernel_flags = flags
        got = ranker.score_candidates(torch.from_numpy(Q).to(DEV), torch.from_numpy(flat).to(DEV),
                                      torch.from_numpy(rowptr).to(DEV)).cpu().numpy()
        assert np.all(got[np.concatenate([index.doclens, index.doclens[::-1]]) == 0] == 0.0)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
        assert rel.max() <= SCORE_RTOL, (dim, flags, rel.max())
    ranker.kernel_flags = 0
    if dim == 128:   # the exhaustive kernel takes the empty documents out and scatters zeros back
        dense = ranker.score_all(torch.from_numpy(Q).to(DEV)).cpu().numpy()
        full = np.stack([O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], pids) for b in range(2)])
        assert np.all(dense[:, index.doclens == 0] == 0.0)
        assert (np.abs(dense - full) / np.maximum(np.abs(full), 1.0)).max() <= SCORE_RTOL
        tp, ts = ranker.rank_exhaustive(torch.from_numpy(Q), k=25)
        want_p, want_s = O.topk_desc(full[0], pids, 25)
        assert np.abs(ts[0].cpu().numpy() - want_s).max() <= SCORE_RTOL * np.maximum(np.abs(want_s), 1.0).max()
    p, s = ranker.rank_forward(torch.from_numpy(Q[0]).unsqueeze(0).permute(0, 2, 1), pids.tolist(), depth=None)
    assert sorted(p) == pids.tolist() and all(a >= b for a, b in zip(s, s[1:]))
    assert all(sc == 0.0 for pid, sc in zip(p, s) if index.doclens[pid] == 0)
