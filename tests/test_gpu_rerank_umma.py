"""The tcgen05 / TMEM rerank kernel (CBK_FLAG_RERANK_TCGEN05) against the oracle and the mma.sync kernel."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as O
from parity_utils import SCORE_RTOL

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
DEV = torch.device("cuda", 0)


def _ranker(index, dt):
    from colbert_b200 import _lib
    from colbert_b200.ranking import ColbertRanker
    emb = torch.from_numpy(index.emb).to(dt)
    r = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
    r.kernel_flags |= _lib.CBK_FLAG_RERANK_TCGEN05
    return r, emb


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("q_len", [32, 8])
def test_umma_rerank_matches_oracle(dt, q_len):
    from colbert_b200 import synthetic
    rng = np.random.default_rng(7)
    doclens = np.concatenate([rng.integers(1, 181, size=2500), np.ones(200, np.int64), rng.integers(129, 513, size=60)])
    rng.shuffle(doclens)
    index = synthetic.make_index(4242, len(doclens), dim=128, doclens=doclens.astype(np.int64))
    ranker, emb = _ranker(index, dt)
    B, n = 6, 500
    Q = synthetic.make_queries(11, B, q_len, 128)
    cand = synthetic.make_candidates(12, B, index.num_docs, n)
    rowptr = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=DEV)
    got = ranker.score_candidates(torch.from_numpy(Q).to(DEV), torch.from_numpy(cand).reshape(-1).to(DEV), rowptr)
    got = got.cpu().numpy().reshape(B, n)
    store, pf = O.pad_store(emb.float().numpy()), O.doclens_pfxsum(index.doclens)
    worst = 0.0
    for b in range(B):
        ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], cand[b])
        worst = max(worst, float((np.abs(got[b] - ref) / np.maximum(np.abs(ref), 1.0)).max()))
    print(f"[tcgen05 {dt} q_len={q_len}] worst relative error {worst:.3e}")
    assert worst <= SCORE_RTOL


def test_umma_rerank_ragged_queries_and_foreign_pids():
    """Many tiny queries (the query buffer changes constantly), empty queries, out-of-range pids."""
    from colbert_b200 import _lib, synthetic
    index = synthetic.make_index(99, 800, dim=128, lo=1, hi=150)
    ranker, emb = _ranker(index, torch.float16)
    rng = np.random.default_rng(5)
    lens = [0, 1, 1, 2, 130, 0, 3, 65, 700, 1, 0, 5]
    Q = synthetic.make_queries(8, len(lens), 32, 128)
    cands = [rng.integers(0, index.num_docs, size=l) for l in lens]
    cands[4][:3] = [-1, 800, 5_000_000]
    flat = np.concatenate(cands).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    got = ranker.score_candidates(torch.from_numpy(Q).to(DEV), torch.from_numpy(flat).to(DEV),
                                  torch.from_numpy(rowptr).to(DEV)).cpu().numpy()
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    for b, l in enumerate(lens):
        seg = got[rowptr[b]: rowptr[b + 1]]
        for i, pid in enumerate(cands[b]):
            if pid < 0 or pid >= index.num_docs:
                assert np.isnan(seg[i])
            else:
                ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], [pid])[0]
                assert abs(seg[i] - ref) <= SCORE_RTOL * max(1.0, abs(ref))


def test_umma_rerank_equals_mma_sync_kernel_at_scale():
    from colbert_b200 import _lib, synthetic
    index = synthetic.make_index(2025, 60_000, dim=128, lo=1, hi=180)
    ranker, emb = _ranker(index, torch.bfloat16)
    B, n = 128, 1000
    Q = torch.from_numpy(synthetic.make_queries(1, B, 32, 128)).to(DEV)
    cand = torch.from_numpy(synthetic.make_candidates(2, B, index.num_docs, n)).to(DEV).reshape(-1)
    rowptr = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=DEV)
    s_umma = ranker.score_candidates(Q, cand, rowptr)
    ranker.kernel_flags &= ~_lib.CBK_FLAG_RERANK_TCGEN05
    s_mma = ranker.score_candidates(Q, cand, rowptr)
    assert torch.allclose(s_umma, s_mma, rtol=1e-3, atol=1e-3)
