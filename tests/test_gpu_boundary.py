"""The reference's plugin / attribute seams on the GPU ranker: ``views`` (colbert_ranker.py:45-51), the injected
``model=`` scorer (colbert_ranker.py:28,111), ``ColbertIndex``'s own constructor signature (colbert_ranker.py:141),
shard-local rankers (pid_base, corpus-wide strides) on the reference-shaped calls.  Needs a B200: ``pytest -m gpu``."""
import os

import numpy as np
import pytest
import torch

from golden_cases import CASES, build_case
from oracle import maxsim_oracle as O
from parity_utils import SCORE_RTOL, check_topk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from colbert_b200 import _lib
    assert _lib.load().cbk_device_supported(0) == 1, "not an sm_100 device"
    return torch.device("cuda", 0)


def make_ranker(index, dev, store_dtype=torch.float16, model=None):
    from colbert_b200.ranking import ColbertRanker
    return ColbertRanker.from_tensors(torch.from_numpy(index.emb), index.doclens.tolist(), device=dev,
                                      store_dtype=store_dtype, model=model)


def test_views_are_the_reference_stride_views(dev):
    """``ranker.views[g][o]`` = store rows o .. o+stride_g-1 (zero copy), one view per stride, as
    reference colbert_ranker.py:45-51 builds them with as_strided."""
    index, _, _ = build_case(CASES[0])
    ranker = make_ranker(index, dev)
    store_np = O.pad_store(index.emb)
    assert len(ranker.views) == len(ranker.strides)
    for view, stride in zip(ranker.views, ranker.strides):
        assert tuple(view.shape) == (ranker.tensor.size(0) - stride + 1, stride, ranker.dim)
        assert view.stride() == (ranker.dim, ranker.dim, 1)
        assert view.data_ptr() == ranker.tensor.data_ptr()                           # a view, not a copy
        expect = torch.as_strided(torch.from_numpy(store_np), tuple(view.shape), (ranker.dim, ranker.dim, 1))
        for o in (0, 1, 777, view.size(0) - 1):                                      # incl. the over-read into the zero tail
            assert torch.equal(view[o].cpu(), expect[o])
    # index_select through a view returns what the reference's gather returns (colbert_ranker.py:105)
    pf = torch.from_numpy(O.doclens_pfxsum(index.doclens))
    offs = pf[torch.tensor([0, 5, 17])].to(dev)
    got = torch.index_select(ranker.views[1], 0, offs).cpu().numpy()
    ref, _ = O.gather_rows(store_np, index.doclens, O.doclens_pfxsum(index.doclens), [0, 5, 17], ranker.strides[1])
    assert np.array_equal(got.astype(np.float32), ref)
    # views follow the stride list when a shard is given the corpus-wide strides
    ranker.strides = [10, 20]
    assert [v.size(1) for v in ranker.views] == [10, 20]


class EinsumScorer:
    """What a caller of the reference would inject as ``model=``: the reference's formula in plain torch ops
    (BaseModel.py:39-46), written here independently of the package under test."""
    calls = 0

    @staticmethod
    def score(Q, D, q_mask, d_mask):
        EinsumScorer.calls += 1
        D = D * d_mask[..., None]
        Q = Q * q_mask[..., None]
        return torch.einsum("qmh,dnh->qdmn", Q, D).max(-1)[0].sum(-1)


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[2]], ids=lambda c: c["name"])
def test_injected_model_is_called_like_the_reference_calls_it(dev, golden_dir, case):
    g = np.load(os.path.join(golden_dir, f"rank_{case['name']}.npz"))
    index, queries, cands = build_case(case)
    EinsumScorer.calls = 0
    ranker = make_ranker(index, dev, model=EinsumScorer)
    fused = make_ranker(index, dev)
    for qi, (Q, pids) in enumerate(zip(queries, cands)):
        Qt = torch.from_numpy(Q).unsqueeze(0).permute(0, 2, 1)
        for dname, depth in case["depths"]:
            p, s = ranker.rank_forward(Qt, [int(x) for x in pids], depth=depth)
            # fp32 einsum on the same gathered rows: this is the reference's arithmetic, so 1e-5, not 1e-3
            check_topk(p, s, g[f"q{qi}_{dname}_pids"], g[f"q{qi}_{dname}_scores"], 1e-5, g[f"q{qi}_all_pids"],
                       g[f"q{qi}_all_scores"])
            p2, s2 = fused.rank_forward(Qt, [int(x) for x in pids], depth=depth)
            check_topk(p2, s2, p, s, SCORE_RTOL, *ranker.rank_forward(Qt, [int(x) for x in pids], depth=None))
        if case.get("output_D"):
            p, D, M = ranker.rank_forward(Qt, [int(x) for x in pids], depth=case["output_D"], output_D_embedding=True)
            assert p == g[f"q{qi}_D_pids"].tolist()
            assert np.array_equal(D.cpu().numpy(), g[f"q{qi}_D_rows"].astype(np.float32))
            assert np.array_equal(M.cpu().numpy(), g[f"q{qi}_D_mask"])
    assert EinsumScorer.calls > 0
    with pytest.raises(TypeError):
        make_ranker(index, dev, model=object())


def test_builtin_basemodel_as_model_runs_the_fused_kernel(dev):
    from colbert_b200.modeling.BaseModel import BaseModel
    index, queries, cands = build_case(CASES[0])
    a, b = make_ranker(index, dev, model=BaseModel), make_ranker(index, dev)
    assert a._model_is_builtin and make_ranker(index, dev, model=BaseModel())._model_is_builtin
    Qt = torch.from_numpy(queries[0]).unsqueeze(0).permute(0, 2, 1)
    assert a.rank_forward(Qt, cands[0].tolist(), depth=10) == b.rank_forward(Qt, cands[0].tolist(), depth=10)


def test_shard_local_ranker_on_the_reference_shaped_calls(dev, tmp_path):
    """A shard ranker (from_flat with a pid range) adopted by ShardedColbertRanker gets the corpus-wide strides and a
    pid_base; rank_forward with a HOST query (one library call), with a device query, and with output_D_embedding
    must all use them: global pids in, corpus-wide zero-floor rule, rows of the right documents out."""
    from colbert_b200 import synthetic
    from colbert_b200.indexing.flat_store import convert_index
    from colbert_b200.ranking import ColbertRanker
    from colbert_b200.sharding import ShardedColbertRanker
    rng = np.random.default_rng(5)
    # shard 1 holds only short documents: its own strides differ from (and are fewer than) the corpus-wide ones
    doclens = np.concatenate([rng.integers(40, 181, size=300), rng.integers(1, 9, size=200)]).astype(np.int64)
    index = synthetic.make_index(91, 500, dim=128, doclens=doclens, num_parts=2)
    src, flat = tmp_path / "idx", tmp_path / "flat"
    src.mkdir()
    synthetic.write_index(index, str(src))
    convert_index(str(src), str(flat))
    gstrides = O.compute_strides(index.doclens)
    local = ColbertRanker.from_flat(str(flat), device=dev, pid_lo=300, pid_hi=500)
    assert local.strides != gstrides and len(local.strides) <= len(gstrides)
    ShardedColbertRanker(local, 300, gstrides)
    assert local.strides == gstrides and list(local._strides_c)[: len(gstrides)] == gstrides
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    Q = synthetic.make_queries(92, 1, 32, 128)[0]
    Q = (-np.abs(Q)).astype(np.float32)                     # negative similarities against ...
    emb = np.abs(index.emb.astype(np.float32)).astype(np.float16)   # ... a positive store: the zero floor decides
    index = synthetic.SynthIndex(emb=emb, doclens=index.doclens, part_sizes=index.part_sizes)
    src2, flat2 = tmp_path / "idx2", tmp_path / "flat2"
    src2.mkdir()
    synthetic.write_index(index, str(src2))
    convert_index(str(src2), str(flat2))
    local = ColbertRanker.from_flat(str(flat2), device=dev, pid_lo=300, pid_hi=500)
    ShardedColbertRanker(local, 300, gstrides)
    store = O.pad_store(index.emb)
    pids = np.arange(300, 500, dtype=np.int64)[::-1].copy()
    ref = O.maxsim_exact(store, index.doclens, pf, gstrides, Q, pids)
    stale = O.maxsim_exact(store, index.doclens, pf, O.compute_strides(index.doclens[300:]), Q, pids)
    assert np.abs(ref - stale).max() > 0.1                  # the two stride lists give different answers here
    rp, rs = O.topk_desc(ref, pids, 20)
    fp, fs = O.topk_desc(ref, pids, None)
    Qt_host = torch.from_numpy(Q).unsqueeze(0).permute(0, 2, 1)
    p, s = local.rank_forward(Qt_host, pids.tolist(), depth=20)                      # cbk_rank_forward_host
    check_topk(p, s, rp, rs, SCORE_RTOL, fp, fs)
    p, s = local.rank_forward(Qt_host.to(dev), pids.tolist(), depth=20)              # device query: two launches
    check_topk(p, s, rp, rs, SCORE_RTOL, fp, fs)
    # output_D_embedding on a shard: doclens / offsets are indexed by LOCAL pid
    same_bucket = [int(x) for x in pids if 1 <= index.doclens[x] <= gstrides[0]][:30]
    p, D, M = local.rank_forward(Qt_host, same_bucket, depth=5, output_D_embedding=True)
    for i, pid in enumerate(p):
        rows = store[pf[pid]: pf[pid] + gstrides[0]].astype(np.float32)
        assert np.array_equal(D[i].cpu().numpy(), rows)
        assert M[i].cpu().numpy().sum() == index.doclens[pid]


def test_colbert_index_reference_constructor_and_shard_pids(dev, tmp_path):
    """``ColbertIndex(index_path, faiss_index_path, nprobe, rank)`` (reference colbert_ranker.py:141) builds emb2pid
    from the index directory; around a shard ranker the pids it emits are GLOBAL."""
    from colbert_b200 import synthetic
    from colbert_b200.ranking.colbert_index import ColbertIndex
    from colbert_b200.ranking import ColbertRanker
    index = synthetic.make_index(61, 300, dim=128, lo=1, hi=30, num_parts=2)
    synthetic.write_index(index, str(tmp_path))
    ref_map = np.repeat(np.arange(index.num_docs), index.doclens)

    def searcher(Q_rows, depth):                                  # deterministic stand-in for the ANN search
        base = torch.arange(Q_rows.size(0), device=Q_rows.device).unsqueeze(1) * 7
        return (base + torch.arange(depth, device=Q_rows.device).unsqueeze(0) * 13) % index.num_tokens

    ci = ColbertIndex(str(tmp_path), searcher, 32, rank=0)
    assert np.array_equal(ci.emb2pid.cpu().numpy(), ref_map.astype(np.int32))
    Q = torch.from_numpy(synthetic.make_queries(62, 2, 4, 128)).to(dev)
    got = ci.retrieve(8, Q)
    ids = searcher(Q.reshape(8, 128), 8).reshape(2, 32).cpu().numpy()
    assert got == [sorted(set(ref_map[ids[b]].tolist())) for b in range(2)]
    with pytest.raises(RuntimeError):
        ColbertIndex(str(tmp_path), "/nonexistent/ivfpq.faiss", 32)        # faiss is not part of this library
    # shard ranker: rows are local, pids come out global
    pf = O.doclens_pfxsum(index.doclens)
    lo = 100
    shard = ColbertRanker.from_tensors(torch.from_numpy(index.emb[pf[lo]:]), index.doclens[lo:].tolist(), device=dev)
    shard.pid_base = lo
    ci2 = ColbertIndex(shard, lambda Qr, d: searcher(Qr, d) % int(pf[-1] - pf[lo]), 32)
    got2 = ci2.retrieve(8, Q)
    ids2 = (ids % int(pf[-1] - pf[lo]))
    local_map = np.repeat(np.arange(index.num_docs - lo), index.doclens[lo:])
    assert got2 == [sorted(set((local_map[ids2[b]] + lo).tolist())) for b in range(2)]


def test_sharded_exhaustive_pads_a_shard_smaller_than_k(dev):
    """world = 1 stand-in for a tiny shard: k_local = 30 < k = 50, the list is padded with key 0 before the merge."""
    from colbert_b200 import synthetic
    from colbert_b200.sharding import ShardedColbertRanker
    index = synthetic.make_index(83, 30, dim=128, lo=1, hi=40)
    sharded = ShardedColbertRanker.from_global_tensors(torch.from_numpy(index.emb), index.doclens, dev)
    Q = synthetic.make_queries(84, 3, 32, 128)
    pids, scores = sharded.rank_exhaustive(torch.from_numpy(Q), k=50)
    assert tuple(pids.shape) == (3, 50)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    allp = np.arange(30, dtype=np.int64)
    for b in range(3):
        ref = O.maxsim_exact(store, index.doclens, pf, sharded.strides, Q[b], allp)
        rp, rs = O.topk_desc(ref, allp, 30)
        check_topk(pids[b, :30].cpu().numpy(), scores[b, :30].cpu().numpy(), rp, rs, SCORE_RTOL)
        assert (pids[b, 30:] == -1).all() and torch.isneginf(scores[b, 30:]).all()
