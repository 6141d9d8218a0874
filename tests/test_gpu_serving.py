"""Batched serving entry: per-query q_lens, ragged candidate lists, the fixed-doclen (multi-view)
fast path, the stream pipeline and the CUDA-graph low-batch path — all against the oracle.  Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as O
from parity_utils import SCORE_RTOL, check_topk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from colbert_b200 import _lib
    assert _lib.load().cbk_device_supported(0) == 1, "not an sm_100 device"
    return torch.device("cuda", 0)


def _ranker(index, dev, store_dtype=torch.float16):
    from colbert_b200.ranking import ColbertRanker
    return ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=dev, store_dtype=store_dtype)


def _oracle_scores(index, strides, Q, ql, cands):
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    return O.maxsim_exact(store, index.doclens, pf, strides, Q[:ql], cands)


@pytest.mark.parametrize("dim,flag", [(128, 0), (128, "tcgen05"), (256, 0), (96, 0), (768, 0)],
                         ids=["mma128", "tcgen05", "wide256", "generic96", "wide768"])
def test_per_query_q_lens(dev, dim, flag):
    """Queries of different lengths in one batch: rows past q_lens[b] are padding and must not matter, whatever they hold
    (here: large garbage).  Equals the oracle on the truncated query — the reference strips the padding before it scores
    (dense_server_client.py:44-46)."""
    from colbert_b200 import _lib, synthetic
    index = synthetic.make_index(301, 1500, dim=dim, lo=1, hi=90)
    ranker = _ranker(index, dev)
    if flag == "tcgen05":
        ranker.kernel_flags |= _lib.CBK_FLAG_RERANK_TCGEN05
    B, n = 9, 120
    q_lens = np.array([32, 1, 5, 16, 17, 31, 8, 32, 3], dtype=np.int32)
    Q = synthetic.make_queries(302, B, 32, dim)
    Qpad = Q.copy()
    for b in range(B):
        Qpad[b, q_lens[b]:] = 7.5                                    # garbage in the padding rows
    cand = synthetic.make_candidates(303, B, index.num_docs, n)
    pids, scores = ranker.rank_forward_batch(torch.from_numpy(Qpad), torch.from_numpy(cand), depth=None,
                                             q_lens=torch.from_numpy(q_lens))
    for b in range(B):
        ref = _oracle_scores(index, ranker.strides, Q[b], int(q_lens[b]), cand[b])
        rp, rs = O.topk_desc(ref, cand[b], None)
        check_topk(pids[b].cpu().numpy(), scores[b].cpu().numpy(), rp, rs, SCORE_RTOL)
    # without q_lens the garbage counts: the answers must differ (the argument is really used)
    _, s2 = ranker.rank_forward_batch(torch.from_numpy(Qpad), torch.from_numpy(cand), depth=None)
    assert not torch.allclose(s2[1], scores[1])


def test_q_lens_with_queries_longer_than_32_rows(dev):
    from colbert_b200 import synthetic
    index = synthetic.make_index(311, 800, dim=128, lo=1, hi=60)
    ranker = _ranker(index, dev)
    B, n, L = 4, 80, 48
    q_lens = np.array([48, 33, 7, 40], dtype=np.int32)
    Q = synthetic.make_queries(312, B, L, 128)
    Qpad = Q.copy()
    for b in range(B):
        Qpad[b, q_lens[b]:] = -3.0
    cand = synthetic.make_candidates(313, B, index.num_docs, n)
    pids, scores = ranker.rank_forward_batch(torch.from_numpy(Qpad), torch.from_numpy(cand), depth=None,
                                             q_lens=torch.from_numpy(q_lens))
    for b in range(B):
        ref = _oracle_scores(index, ranker.strides, Q[b], int(q_lens[b]), cand[b])
        rp, rs = O.topk_desc(ref, cand[b], None)
        check_topk(pids[b].cpu().numpy(), scores[b].cpu().numpy(), rp, rs, SCORE_RTOL)


@pytest.mark.parametrize("store_dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("d_view,q_view", [(8, 8), (16, 16), (4, 8), (12, 32)])
def test_fixed_doclen_fast_path(dev, d_view, q_view, store_dtype):
    """Multi-view index (every document = d_view rows): the ranker recognises it, the kernel computes offsets as
    pid * d_view without touching pfxsum / doclens — bit-identical to the looked-up path, and equal to the oracle."""
    from colbert_b200 import _lib, kernels, synthetic
    index = synthetic.make_index(321, 3000, dim=128, doclen_kind="fixed", fixed=d_view)
    emb = torch.from_numpy(index.emb).to(store_dtype)
    from colbert_b200.ranking import ColbertRanker
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=dev, store_dtype=store_dtype)
    assert ranker.strides == [d_view] and ranker.effective_flags & _lib.CBK_FLAG_FIXED_DOCLEN
    B, n = 6, 500
    Q = synthetic.make_queries(322, B, q_view, 128)
    cand = synthetic.make_candidates(323, B, index.num_docs, n)
    cand[0, :4] = [0, index.num_docs - 1, 1, index.num_docs - 2]
    Qd, cd = torch.from_numpy(Q).to(dev), torch.from_numpy(cand).reshape(-1).to(dev)
    rowptr = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=dev)
    fast = ranker.score_candidates(Qd, cd, rowptr)
    slow = kernels.maxsim_rerank(ranker.tensor, ranker._pfxsum_dev, ranker._doclens_dev, ranker.strides, Qd, cd, rowptr,
                                 flags=ranker.kernel_flags)
    assert torch.equal(fast, slow)
    # the metadata arrays are really not read: poison them
    bad_pf, bad_dl = torch.full_like(ranker._pfxsum_dev, 1 << 40), torch.full_like(ranker._doclens_dev, -7)
    poisoned = kernels.maxsim_rerank(ranker.tensor, bad_pf, bad_dl, ranker.strides, Qd, cd, rowptr, flags=ranker.effective_flags)
    assert torch.equal(fast, poisoned)
    store = O.pad_store(emb.float().numpy())
    pf = O.doclens_pfxsum(index.doclens)
    got = fast.cpu().numpy().reshape(B, n)
    for b in range(B):
        ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], cand[b])
        rel = np.abs(got[b] - ref) / np.maximum(np.abs(ref), 1.0)
        assert rel.max() <= SCORE_RTOL
    # a shard of such an index given corpus-wide strides that are NOT [d_view] must fall back to the looked-up path
    ranker.strides = [d_view, d_view + 5]
    assert not ranker.effective_flags & _lib.CBK_FLAG_FIXED_DOCLEN


def test_pipeline_ragged_lists_and_q_lens(dev):
    from colbert_b200 import synthetic
    from colbert_b200.ranking.pipeline import RerankPipeline
    index = synthetic.make_index(331, 2500, dim=128, lo=1, hi=100)
    ranker = _ranker(index, dev)
    B, nmax = 12, 150
    rng = np.random.default_rng(332)
    lens = rng.integers(0, nmax + 1, size=B)
    lens[3] = 0
    lens[5] = nmax
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = np.concatenate([rng.choice(index.num_docs, size=l, replace=False) for l in lens]).astype(np.int64)
    q_lens = rng.integers(1, 33, size=B).astype(np.int32)
    Q = synthetic.make_queries(333, B, 32, 128)
    Qpad = Q.copy()
    for b in range(B):
        Qpad[b, q_lens[b]:] = 11.0
    pipe = RerankPipeline(ranker, B, 32, nmax, depth=10)
    flat_pin = torch.zeros(B * nmax, dtype=torch.int64).pin_memory()
    flat_pin[: flat.size] = torch.from_numpy(flat)
    for _ in range(3):                                                   # slots are reused
        h = pipe.submit(torch.from_numpy(Qpad).pin_memory(), flat_pin, q_lens=torch.from_numpy(q_lens),
                        cand_rowptr=torch.from_numpy(rp))
        pids, scores = pipe.result(h)
    for b in range(B):
        c = flat[rp[b]: rp[b + 1]]
        k = min(10, c.size)
        if k:
            ref = _oracle_scores(index, ranker.strides, Q[b], int(q_lens[b]), c)
            rp_, rs_ = O.topk_desc(ref, c, k)
            fp, fs = O.topk_desc(ref, c, None)
            check_topk(pids[b, :k].numpy(), scores[b, :k].numpy(), rp_, rs_, SCORE_RTOL, fp, fs)
        assert (pids[b, k:] == -1).all() and torch.isneginf(scores[b, k:]).all()


def test_graphed_rerank_matches_direct_call(dev):
    """CUDA-graph replay of H2D -> MaxSim -> top-k -> D2H for small batches: same answers as rank_forward_batch, for
    every batch size / list length up to the captured shape, call after call."""
    from colbert_b200 import synthetic
    from colbert_b200.ranking.pipeline import GraphedRerank
    index = synthetic.make_index(341, 4000, dim=128, lo=1, hi=120)
    ranker = _ranker(index, dev)
    g = GraphedRerank(ranker, batch=8, q_len=32, n_cand=1000, depth=10)
    rng = np.random.default_rng(342)
    for it, (b, n, ql) in enumerate([(1, 1000, 32), (8, 1000, 32), (3, 417, 20), (8, 64, 1), (1, 1000, 32)]):
        Q = synthetic.make_queries(350 + it, b, ql, 128)
        cand = synthetic.make_candidates(360 + it, b, index.num_docs, n)
        q_lens = rng.integers(1, ql + 1, size=b).astype(np.int32)
        p, s = g(torch.from_numpy(Q), torch.from_numpy(cand), q_lens=torch.from_numpy(q_lens))
        rp, rs = ranker.rank_forward_batch(torch.from_numpy(Q), torch.from_numpy(cand), depth=10, q_lens=torch.from_numpy(q_lens))
        assert torch.equal(p, rp.cpu()) and torch.equal(s, rs.cpu())
        ref = _oracle_scores(index, ranker.strides, Q[0], int(q_lens[0]), cand[0])
        op, os_ = O.topk_desc(ref, cand[0], 10)
        fp, fs = O.topk_desc(ref, cand[0], None)
        check_topk(p[0].numpy(), s[0].numpy(), op, os_, SCORE_RTOL, fp, fs)
