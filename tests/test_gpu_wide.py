"""Tensor-core rerank kernels for wide embeddings (dim a multiple of 64 other than 128; the author's configuration uses
768) against the oracle and against the generic CUDA-core kernel: the tcgen05 streaming kernel (dim 192 … 1024, the default
there, ragged and fixed-length multi-view stores), the K-split mma.sync kernel (dim 64, or
CBK_FLAG_RERANK_KSPLIT)."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as O
from parity_utils import SCORE_RTOL, check_topk

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("dim", [64, 192, 256, 320, 384, 448, 512, 768, 832, 1024])   # 320 / 448 / 832: uneven k-step split
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_wide_rerank_matches_oracle_and_generic(dim, dt):
    from colbert_b200 import _lib, synthetic
    from colbert_b200.ranking import ColbertRanker
    rng = np.random.default_rng(dim)
    doclens = np.concatenate([rng.integers(1, 200, size=250), [0, 0, 1, 15, 16, 17, 31, 32, 33, 400]]).astype(np.int64)
    index = synthetic.make_index(900 + dim, len(doclens), dim=dim, doclens=doclens)
    emb = torch.from_numpy(index.emb).to(dt)
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
    store, pf = O.pad_store(emb.float().numpy()), O.doclens_pfxsum(index.doclens)
    for q_len in (32, 7):
        n_q = 5
        Q = synthetic.make_queries(901 + q_len, n_q, q_len, dim)
        lens = [0, 1, 40, 97, 260]
        cands = [rng.integers(0, index.num_docs, size=l).astype(np.int64) for l in lens]
        cands[4][:10] = np.arange(250, 260)                                        # the edge-length documents
        flat = np.concatenate(cands)
        rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        ref = np.concatenate([O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], cands[b]) if lens[b] else
                              np.zeros(0, np.float32) for b in range(n_q)])
        got = {}
        for name, flags in (("wide", 0), ("ksplit", _lib.CBK_FLAG_RERANK_KSPLIT), ("generic", _lib.CBK_FLAG_RERANK_GENERIC)):
            ranker.kernel_flags = flags
            got[name] = ranker.score_candidates(torch.from_numpy(Q).to(DEV), torch.from_numpy(flat).to(DEV),
                                                torch.from_numpy(rowptr).to(DEV)).cpu().numpy()
            rel = np.abs(got[name] - ref) / np.maximum(np.abs(ref), 1.0)
            assert rel.max() <= SCORE_RTOL, (dim, q_len, name, rel.max())
        assert np.abs(got["wide"] - got["generic"]).max() <= SCORE_RTOL * max(1.0, np.abs(ref).max())
        assert np.abs(got["ksplit"] - got["generic"]).max() <= SCORE_RTOL * max(1.0, np.abs(ref).max())
    ranker.kernel_flags = 0


def test_wide_rerank_foreign_pids_and_rank_forward():
    """dim 768 end to end through rank_forward, plus out-of-range pids (NaN, or -inf on a shard)."""
    from colbert_b200 import _lib, synthetic
    from colbert_b200.ranking import ColbertRanker
    dim = 768
    index = synthetic.make_index(77, 600, dim=dim, lo=1, hi=120)
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=DEV)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    Q = synthetic.make_queries(78, 1, 32, dim)[0]
    pids = np.random.default_rng(79).permutation(600)[:300].astype(np.int64)
    p, s = ranker.rank_forward(torch.from_numpy(Q).unsqueeze(0).permute(0, 2, 1), pids.tolist(), depth=20)
    ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q, pids)
    rp, rs = O.topk_desc(ref, pids, 20)
    check_topk(p, s, rp, rs, SCORE_RTOL, *O.topk_desc(ref, pids, None))
    bad = torch.tensor([5, 600, -1, 7], dtype=torch.int64, device=DEV)
    rp2 = torch.tensor([0, 4], dtype=torch.int64, device=DEV)
    sc = ranker.score_candidates(torch.from_numpy(Q[None]).to(DEV), bad, rp2).cpu().numpy()
    assert np.isnan(sc[1]) and np.isnan(sc[2]) and np.isfinite(sc[0]) and np.isfinite(sc[3])
    ranker.kernel_flags = _lib.CBK_FLAG_SKIP_FOREIGN_PIDS
    sc = ranker.score_candidates(torch.from_numpy(Q[None]).to(DEV), bad, rp2).cpu().numpy()
    assert np.isneginf(sc[1]) and np.isneginf(sc[2])


@pytest.mark.parametrize("dim", [256, 512, 768, 1024])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_multiview_16_rows_wide_streaming_kernel(dim, dt):
    """the author's operating point (16 view embeddings per document and per query, un-projected width): a fixed-length
    store (CBK_FLAG_FIXED_DOCLEN: row = pid * 16, no metadata lookups) through the tcgen05 streaming kernel against the oracle and the generic kernel — ragged lists, empty lists, short queries, per-query
    q_lens, pids outside the store."""
    from colbert_b200 import _lib, synthetic
    from colbert_b200.ranking import ColbertRanker
    rng = np.random.default_rng(dim + 1)
    n_docs = 700
    doclens = np.full(n_docs, 16, dtype=np.int64)
    index = synthetic.make_index(1300 + dim, n_docs, dim=dim, doclens=doclens)
    emb = torch.from_numpy(index.emb).to(dt)
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
    assert ranker.strides == [16] and ranker.effective_flags & _lib.CBK_FLAG_FIXED_DOCLEN
    store, pf = O.pad_store(emb.float().numpy()), O.doclens_pfxsum(index.doclens)
    for q_len in (16, 5):
        lens = [3, 0, 8, 1000, 17, 64, 1, 300]
        n_q = len(lens)
        Q = synthetic.make_queries(1301 + q_len, n_q, q_len, dim)
        cands = [rng.integers(0, n_docs, size=l).astype(np.int64) for l in lens]
        flat = np.concatenate(cands)
        rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        ref = np.concatenate([O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], cands[b]) if lens[b] else
                              np.zeros(0, np.float32) for b in range(n_q)])
        args = (torch.from_numpy(Q).to(DEV), torch.from_numpy(flat).to(DEV), torch.from_numpy(rowptr).to(DEV))
        launches0 = _lib.launch_count()
        got = ranker.score_candidates(*args).cpu().numpy()
        assert _lib.launch_count() == launches0 + 1
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
        assert rel.max() <= SCORE_RTOL, (dim, q_len, rel.max())
        # bf16 stores: the query enters as bf16 value + bf16 residual, so the only rounding left is the store's own
        if dt == torch.bfloat16:
            assert rel.max() <= 5e-5, rel.max()
        ranker.kernel_flags = _lib.CBK_FLAG_RERANK_GENERIC
        gen = ranker.score_candidates(*args).cpu().numpy()
        ranker.kernel_flags = 0
        assert np.abs(got - gen).max() <= SCORE_RTOL * max(1.0, np.abs(ref).max())
    # per-query real lengths inside a 16-row slot: rows at or past q_lens[q] are padding whatever they hold
    q_lens = np.array([16, 1, 9, 16, 4, 12, 16, 7], dtype=np.int32)
    Qpad = synthetic.make_queries(1400, n_q, 16, dim)
    got = ranker.score_candidates(torch.from_numpy(Qpad).to(DEV), torch.from_numpy(flat).to(DEV), torch.from_numpy(rowptr).to(DEV),
                                  q_lens=torch.from_numpy(q_lens).to(DEV)).cpu().numpy()
    ref = np.concatenate([O.maxsim_exact(store, index.doclens, pf, ranker.strides, Qpad[b][:q_lens[b]], cands[b]) if lens[b] else
                          np.zeros(0, np.float32) for b in range(n_q)])
    assert (np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)).max() <= SCORE_RTOL
    # pids outside the store: NaN, or -inf on a shard
    bad = torch.tensor([5, n_docs, -1, 7, 8, 9, 10, 11, 12], dtype=torch.int64, device=DEV)
    rp2 = torch.tensor([0, 9], dtype=torch.int64, device=DEV)
    sc = ranker.score_candidates(torch.from_numpy(Qpad[:1]).to(DEV), bad, rp2).cpu().numpy()
    assert np.isnan(sc[1]) and np.isnan(sc[2]) and np.isfinite(np.delete(sc, [1, 2])).all()
    ranker.kernel_flags = _lib.CBK_FLAG_SKIP_FOREIGN_PIDS
    sc = ranker.score_candidates(torch.from_numpy(Qpad[:1]).to(DEV), bad, rp2).cpu().numpy()
    assert np.isneginf(sc[1]) and np.isneginf(sc[2])
    ranker.kernel_flags = 0


def test_multiview_wide_many_ctas_and_query_changes():
    """enough candidates for every CTA to own a range that spans several queries (query region swapped under the ring)"""
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    dim, n_docs, n_q, n_c = 768, 4000, 600, 37
    index = synthetic.make_index(1500, n_docs, dim=dim, doclens=np.full(n_docs, 16, dtype=np.int64))
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=DEV)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    Q = synthetic.make_queries(1501, n_q, 16, dim)
    cand = np.random.default_rng(1502).integers(0, n_docs, size=(n_q, n_c)).astype(np.int64)
    rowptr = torch.arange(0, (n_q + 1) * n_c, n_c, dtype=torch.int64, device=DEV)
    got = ranker.score_candidates(torch.from_numpy(Q).to(DEV), torch.from_numpy(cand.reshape(-1)).to(DEV), rowptr).cpu().numpy()
    for b in range(0, n_q, 7):
        ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], cand[b])
        assert (np.abs(got[b * n_c:(b + 1) * n_c] - ref) / np.maximum(np.abs(ref), 1.0)).max() <= SCORE_RTOL, b


def test_wide_stream_long_documents_and_many_queries():
    """documents longer than a tile (> 128 rows: the running maximum crosses tiles), every CTA owning a range that spans
    several queries, lists that end in the middle of a tile, q_len 20 (both halves of the query rows, the second partial)"""
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    dim, n_docs, n_q, n_c = 512, 1500, 333, 29
    rng = np.random.default_rng(1600)
    doclens = rng.integers(1, 400, size=n_docs).astype(np.int64)
    index = synthetic.make_index(1601, n_docs, dim=dim, doclens=doclens)
    for dt in (torch.float16, torch.bfloat16):
        emb = torch.from_numpy(index.emb).to(dt)
        ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
        store, pf = O.pad_store(emb.float().numpy()), O.doclens_pfxsum(index.doclens)
        Q = synthetic.make_queries(1602, n_q, 20, dim)
        cand = rng.integers(0, n_docs, size=(n_q, n_c)).astype(np.int64)
        rowptr = torch.arange(0, (n_q + 1) * n_c, n_c, dtype=torch.int64, device=DEV)
        got = ranker.score_candidates(torch.from_numpy(Q).to(DEV), torch.from_numpy(cand.reshape(-1)).to(DEV), rowptr).cpu().numpy()
        for b in range(0, n_q, 11):
            ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], cand[b])
            assert (np.abs(got[b * n_c:(b + 1) * n_c] - ref) / np.maximum(np.abs(ref), 1.0)).max() <= SCORE_RTOL, (dt, b)


def test_wide_stream_scores_are_bit_reproducible():
    """q_len > 16: the two halves of the query rows are ADDED into a zeroed score (two commutative additions) — the result
    must not depend on which epilogue team gets there first"""
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    dim, n_docs, n_q, n_c = 384, 2000, 64, 200
    rng = np.random.default_rng(1700)
    doclens = rng.integers(1, 181, size=n_docs).astype(np.int64)
    index = synthetic.make_index(1701, n_docs, dim=dim, doclens=doclens)
    emb = torch.from_numpy(index.emb).to(torch.bfloat16)
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=torch.bfloat16)
    Q = torch.from_numpy(synthetic.make_queries(1702, n_q, 32, dim)).to(DEV)
    cand = torch.from_numpy(rng.integers(0, n_docs, size=n_q * n_c).astype(np.int64)).to(DEV)
    rowptr = torch.arange(0, (n_q + 1) * n_c, n_c, dtype=torch.int64, device=DEV)
    a = ranker.score_candidates(Q, cand, rowptr).clone()
    for _ in range(3):
        assert torch.equal(a, ranker.score_candidates(Q, cand, rowptr))
