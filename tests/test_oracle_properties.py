"""Property tests of the oracle (CPU, hypothesis): the faithful op-sequence restatement of rank_forward and the
exact-doclen + zero-floor formulation the CUDA kernels implement agree on arbitrary small indexes; sharding plans
partition the corpus; packed keys round-trip and order like (score desc, pid asc)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from colbert_b200.sharding import plan_shards
from oracle import maxsim_oracle as O


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(4, 40), st.integers(1, 24), st.integers(1, 8))
def test_faithful_and_exact_formulations_agree(seed, n_docs, max_len, q_len):
    rng = np.random.default_rng(seed)
    doclens = rng.integers(1, max_len + 1, size=n_docs).astype(np.int64)
    emb = rng.standard_normal((int(doclens.sum()), 16)).astype(np.float16)       # signs vary: the floor matters
    store, pf = O.pad_store(emb), O.doclens_pfxsum(doclens)
    strides = O.compute_strides(doclens)
    Q = rng.standard_normal((q_len, 16)).astype(np.float32)
    pids = rng.permutation(n_docs)[: max(1, n_docs // 2)]
    _, _, all_scores = O.rank_forward(store, doclens, pf, strides, np.transpose(Q[None], (0, 2, 1)), pids, depth=None,
                                      return_all_scores=True)
    exact = O.maxsim_exact(store, doclens, pf, strides, Q, pids)
    np.testing.assert_allclose(exact, all_scores, rtol=1e-5, atol=1e-5)


@settings(max_examples=50, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(1, 300), st.integers(1, 9))
def test_plan_shards_partitions_the_corpus(seed, n_docs, world):
    rng = np.random.default_rng(seed)
    dl = torch.from_numpy(rng.integers(1, 50, size=n_docs))
    pf = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(dl, 0)])
    b = plan_shards(pf, world)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] == n_docs and all(b[i] <= b[i + 1] for i in range(world))


@settings(max_examples=50, deadline=None)
@given(st.lists(st.tuples(st.floats(-100, 100, width=32), st.integers(0, 2**32 - 2)), min_size=1, max_size=60, unique_by=lambda t: t[1]))
def test_packed_keys_order_and_roundtrip(pairs):
    scores = np.array([p[0] for p in pairs], dtype=np.float32)
    pids = np.array([p[1] for p in pairs], dtype=np.int64)
    keys = O.pack_keys(scores, pids)
    s2, p2 = O.unpack_keys(keys)
    assert np.array_equal(p2, pids) and np.array_equal(s2, scores + np.float32(0.0))
    order = np.argsort(keys)[::-1]
    ref_p, ref_s = O.topk_desc(scores, pids, None)
    assert np.array_equal(pids[order], ref_p)


def test_empty_documents_score_zero_in_both_formulations():
    """A document of length 0: the reference gathers the NEXT document's rows under an all-false mask, so every
    query token's maximum is 0 and the score is 0 (colbert_ranker.py:105-109, BaseModel.py:41-46); the exact-length
    formulation reaches the same value through the zero floor (0 is not a stride)."""
    from colbert_b200 import synthetic
    index = synthetic.make_index(77, 300, dim=32, lo=0, hi=12)
    assert 0 < int((index.doclens == 0).sum()) < 70
    strides = O.compute_strides(index.doclens)
    assert 0 not in strides
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    Q = synthetic.make_queries(78, 1, 8, 32)[0]
    pids = np.arange(300)
    exact = O.maxsim_exact(store, index.doclens, pf, strides, Q, pids)
    assert np.all(exact[index.doclens == 0] == 0.0)
    _, _, ref = O.rank_forward(store, index.doclens, pf, strides, Q.T[None], pids, depth=None, return_all_scores=True)
    np.testing.assert_allclose(exact, ref, rtol=1e-5, atol=1e-5)


def test_score_allpairs_grad_is_the_derivative_of_score_allpairs():
    """central differences of the oracle's forward against the oracle's backward (away from arg-max ties the score is
    piecewise linear in Q and D, so the difference quotient is exact up to rounding); linear in the upstream gradient W"""
    import numpy as np
    from oracle import maxsim_oracle as O
    rng = np.random.default_rng(5)
    nq, m, nd, n, h = 2, 3, 3, 4, 5
    Q = rng.standard_normal((nq, m, h))
    D = rng.standard_normal((nd, n, h))
    qm = np.array([[1, 1, 0], [1, 1, 1]])
    dm = np.array([[1, 1, 1, 0], [1, 0, 0, 0], [1, 1, 1, 1]])
    W = rng.standard_normal((nq, nd)).astype(np.float32)
    _, dQ, dD, _ = O.score_allpairs_grad(Q, D, qm, dm, W)

    def loss(Qx, Dx):
        Qm, Dm = Qx * qm[..., None], Dx * dm[..., None]
        sim = np.einsum("qmh,dnh->qdmn", Qm, Dm)
        return float((sim.max(-1).sum(-1) * W).sum())

    eps = 1e-6
    for idx in [(0, 0, 1), (1, 2, 4), (0, 2, 0)]:
        Qp, Qn = Q.copy(), Q.copy()
        Qp[idx] += eps
        Qn[idx] -= eps
        assert abs((loss(Qp, D) - loss(Qn, D)) / (2 * eps) - dQ[idx]) < 1e-4
    for idx in [(0, 0, 0), (1, 0, 3), (2, 3, 2), (0, 3, 1)]:
        Dp, Dn = D.copy(), D.copy()
        Dp[idx] += eps
        Dn[idx] -= eps
        assert abs((loss(Q, Dp) - loss(Q, Dn)) / (2 * eps) - dD[idx]) < 1e-4
    _, dQ2, dD2, _ = O.score_allpairs_grad(Q, D, qm, dm, 2 * W)
    np.testing.assert_allclose(dQ2, 2 * dQ, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dD2, 2 * dD, rtol=1e-6, atol=1e-7)
