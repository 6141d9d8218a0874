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
