"""Query-batched exhaustive scoring (tcgen05 kernel) and the dense top-k against the oracle."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as O
from parity_utils import SCORE_RTOL, check_topk

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _index(seed, doclens):
    from colbert_b200 import synthetic
    return synthetic.make_index(seed, len(doclens), dim=128, doclens=doclens)


def _oracle_scores(index, strides, Q, store_values):
    store, pf = O.pad_store(store_values), O.doclens_pfxsum(index.doclens)
    all_pids = np.arange(index.num_docs)
    return np.stack([O.maxsim_exact(store, index.doclens, pf, strides, Q[b], all_pids) for b in range(Q.shape[0])])


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("n_q,q_len", [(1, 32), (4, 32), (5, 8), (17, 32), (40, 20)])
def test_exhaustive_scores_match_oracle(dt, n_q, q_len):
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    rng = np.random.default_rng(n_q)
    # long documents (span several 128-token tiles), runs of 1-token documents, ordinary ones
    doclens = np.concatenate([rng.integers(1, 181, size=1500), np.ones(300, np.int64), rng.integers(200, 513, size=40),
                              rng.integers(1, 181, size=1200)]).astype(np.int64)
    rng.shuffle(doclens)
    index = _index(1000 + n_q, doclens)
    emb = torch.from_numpy(index.emb).to(dt)
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
    Q = synthetic.make_queries(50 + n_q, n_q, q_len, 128)
    got = ranker.score_all(torch.from_numpy(Q).to(DEV)).cpu().numpy()
    ref = _oracle_scores(index, ranker.strides, Q, emb.float().numpy())
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    print(f"[{dt}, n_q={n_q}, q_len={q_len}] worst relative error {rel.max():.3e}")
    assert got.shape == ref.shape and rel.max() <= SCORE_RTOL


def test_exhaustive_equals_rerank_of_everything():
    """The tcgen05 path and the mma.sync rerank path are two kernels for the same function."""
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    index = synthetic.make_index(4, 2500, dim=128, lo=1, hi=180)
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=DEV)
    Q = torch.from_numpy(synthetic.make_queries(5, 6, 32, 128)).to(DEV)
    dense = ranker.score_all(Q)
    cand = torch.arange(index.num_docs, dtype=torch.int64, device=DEV).repeat(6)
    rowptr = torch.arange(0, 7 * index.num_docs, index.num_docs, dtype=torch.int64, device=DEV)
    rer = ranker.score_candidates(Q, cand, rowptr).reshape(6, -1)
    assert torch.allclose(dense, rer, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("n_docs,k", [(50, 10), (5000, 1000), (40_000, 1000), (300_000, 100)])
def test_topk_dense_matches_oracle(n_docs, k):
    from colbert_b200 import kernels
    rng = np.random.default_rng(n_docs)
    sc = np.round(rng.standard_normal((3, n_docs)), 2).astype(np.float32)      # rounding → exact ties
    ts, tp = kernels.topk_dense(torch.from_numpy(sc).to(DEV), min(k, n_docs), pid_base=1000)
    for b in range(3):
        rp, rs = O.topk_desc(sc[b], np.arange(n_docs, dtype=np.int64) + 1000, min(k, n_docs))
        assert np.array_equal(tp[b].cpu().numpy(), rp) and np.array_equal(ts[b].cpu().numpy(), rs)


@pytest.mark.parametrize("case", ["normal_k8192", "just_over_threshold", "massive_ties", "two_values", "negatives_and_zeros",
                                  "clustered"])
def test_topk_dense_radix_select(case):
    """Rows long enough for the radix-select path (> 32768 scores): exact top-k in the total order (score desc, pid asc)
    including when more than 16384 scores tie with the k-th (those queries fall back to the chunk sort on the device)."""
    from colbert_b200 import kernels
    rng = np.random.default_rng(sum(map(ord, case)))
    n_docs, k, B = 200_000, 1000, 3
    if case == "normal_k8192":
        sc, k = rng.standard_normal((B, n_docs)).astype(np.float32) * 7 + 20, 8192
    elif case == "just_over_threshold":
        n_docs = 32_769
        sc = rng.standard_normal((B, n_docs)).astype(np.float32)
    elif case == "massive_ties":
        sc = rng.integers(0, 3, size=(B, n_docs)).astype(np.float32)             # ~66 k scores tie with the k-th
    elif case == "two_values":
        sc = np.zeros((B, n_docs), dtype=np.float32)
        sc[:, ::1000] = 5.0                                                       # 200 winners, then 199 800 ties at 0
    elif case == "negatives_and_zeros":
        sc = -np.abs(rng.standard_normal((B, n_docs))).astype(np.float32)
        sc[:, rng.integers(0, n_docs, 300)] = 0.0
        sc[:, rng.integers(0, n_docs, 300)] = -0.0                                # -0.0 ties with +0.0, lower pid first
    else:   # clustered: everything inside one binade, differences in the low mantissa bits only
        sc = (16.0 + rng.integers(0, 4000, size=(B, n_docs)) * np.float32(2.0 ** -19)).astype(np.float32)
    dev_sc = torch.from_numpy(sc).to(DEV)
    ts, tp = kernels.topk_dense(dev_sc, k, pid_base=7)
    keys = kernels.topk_dense(dev_sc, k, pid_base=7, as_keys=True)
    pids = np.arange(n_docs, dtype=np.int64) + 7
    for b in range(B):
        rp, rs = O.topk_desc(sc[b] + np.float32(0.0), pids, k)
        assert np.array_equal(tp[b].cpu().numpy(), rp), case
        assert np.array_equal(ts[b].cpu().numpy(), rs + np.float32(0.0)), case
        ks, kp = O.unpack_keys(keys[b].cpu().numpy())
        assert np.array_equal(kp, rp) and np.array_equal(ks, rs + np.float32(0.0)), case


def test_rank_exhaustive_topk():
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    index = synthetic.make_index(8, 20_000, dim=128, lo=20, hi=120)
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=DEV)
    Q = synthetic.make_queries(9, 3, 32, 128)
    pids, scores = ranker.rank_exhaustive(torch.from_numpy(Q), k=100)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    for b in range(3):
        ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], np.arange(index.num_docs))
        rp, rs = O.topk_desc(ref, np.arange(index.num_docs, dtype=np.int64), 100)
        fp, fs = O.topk_desc(ref, np.arange(index.num_docs, dtype=np.int64), None)
        check_topk(pids[b].cpu().numpy(), scores[b].cpu().numpy(), rp, rs, SCORE_RTOL, fp, fs)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_exhaustive_matches_single_store(world):
    """Config 4 layout on one GPU: W shard stores side by side, packed-key lists stacked as the all-gather would,
    merged by the device kernel → identical to the single-store result."""
    from colbert_b200 import kernels, synthetic
    from colbert_b200.ranking import ColbertRanker
    from colbert_b200.sharding import plan_shards
    index = synthetic.make_index(515, 30_000, dim=128, lo=20, hi=120)
    single = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=DEV)
    Q = torch.from_numpy(synthetic.make_queries(516, 5, 32, 128)).to(DEV)
    k = 200
    ref_pids, ref_scores = single.rank_exhaustive(Q, k=k)
    pf = single.doclens_pfxsum
    bounds = plan_shards(pf, world)
    keys = []
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        shard = ColbertRanker.from_tensors(torch.from_numpy(index.emb[int(pf[lo]): int(pf[hi])]),
                                           index.doclens[lo:hi].tolist(), device=DEV)
        shard.strides, shard.pid_base = single.strides, lo
        keys.append(kernels.topk_dense(shard.score_all(Q), k, pid_base=lo, as_keys=True))
    scores, pids = kernels.merge_topk_keys(torch.stack(keys).contiguous(), k)
    assert torch.equal(pids, ref_pids) and torch.equal(scores, ref_scores)


@pytest.mark.parametrize("n_docs", [4, 5, 7, 149])   # < 4 documents: the reference's kthvalue(k=0) raises, so does ours
def test_exhaustive_tiny_corpora(n_docs):
    """Fewer documents than CTA sub-ranges; k larger than the corpus."""
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    index = synthetic.make_index(600 + n_docs, n_docs, dim=128, lo=1, hi=200)
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=DEV)
    Q = synthetic.make_queries(601, 3, 32, 128)
    got = ranker.score_all(torch.from_numpy(Q).to(DEV)).cpu().numpy()
    ref = _oracle_scores(index, ranker.strides, Q, index.emb)
    assert np.abs(got - ref).max() <= SCORE_RTOL * max(1.0, np.abs(ref).max())
    pids, scores = ranker.rank_exhaustive(torch.from_numpy(Q), k=1000)
    assert pids.shape == (3, n_docs) and sorted(pids[0].tolist()) == list(range(n_docs))


@pytest.mark.parametrize("dim,d_view,q_view", [(768, 16, 16), (256, 8, 8), (512, 16, 32)])
def test_exhaustive_on_a_wide_multiview_store(dim, d_view, q_view):
    """a fixed-length fp16 store at the author's un-projected width: score_all / rank_exhaustive run the all-pairs tcgen05
    kernel over the store seen as [n_docs, d_view, dim]"""
    import numpy as np
    import torch
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    from oracle import maxsim_oracle as O
    from parity_utils import SCORE_RTOL, check_topk
    dev = torch.device("cuda", 0)
    n_docs, nq, k = 3001, 5, 50
    index = synthetic.make_index(2100 + dim, n_docs, dim=dim, doclens=np.full(n_docs, d_view, dtype=np.int64))
    ranker = ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=dev)
    Q = synthetic.make_queries(2101, nq, q_view, dim)
    dense = ranker.score_all(torch.from_numpy(Q).to(dev)).cpu().numpy()
    assert dense.shape == (nq, n_docs)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    pids_all = np.arange(n_docs, dtype=np.int64)
    p, s = ranker.rank_exhaustive(torch.from_numpy(Q), k=k)
    for b in range(nq):
        ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], pids_all)
        assert (np.abs(dense[b] - ref) / np.maximum(np.abs(ref), 1.0)).max() <= SCORE_RTOL
        rp, rs = O.topk_desc(ref, pids_all, k)
        check_topk(p[b].cpu().numpy(), s[b].cpu().numpy(), rp, rs, SCORE_RTOL, *O.topk_desc(ref, pids_all, None))


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_exhaustive_on_a_wide_ragged_store(dt):
    """ragged documents at dim 768 (and bf16 stores): score_all falls back to the rerank kernel of that width with every
    document as a candidate of every query"""
    from colbert_b200 import synthetic
    from colbert_b200.ranking import ColbertRanker
    dim, n_docs, nq, k = 768, 1500, 3, 40
    doclens = np.random.default_rng(2200).integers(1, 150, size=n_docs).astype(np.int64)
    index = synthetic.make_index(2201, n_docs, dim=dim, doclens=doclens)
    emb = torch.from_numpy(index.emb).to(dt)
    ranker = ColbertRanker.from_tensors(emb, index.doclens.tolist(), device=DEV, store_dtype=dt)
    Q = synthetic.make_queries(2202, nq, 32, dim)
    store, pf = O.pad_store(emb.float().numpy()), O.doclens_pfxsum(index.doclens)
    pids_all = np.arange(n_docs, dtype=np.int64)
    p, s = ranker.rank_exhaustive(torch.from_numpy(Q), k=k)
    for b in range(nq):
        ref = O.maxsim_exact(store, index.doclens, pf, ranker.strides, Q[b], pids_all)
        rp, rs = O.topk_desc(ref, pids_all, k)
        check_topk(p[b].cpu().numpy(), s[b].cpu().numpy(), rp, rs, SCORE_RTOL, *O.topk_desc(ref, pids_all, None))
