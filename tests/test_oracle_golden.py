"""The oracle against the reference: every golden vector produced by running the UNMODIFIED
reference (tests/golden/make_golden.py) must be reproduced by oracle/maxsim_oracle.py on the same
seeded inputs.  CPU only."""
import os

import numpy as np
import pytest

from golden_cases import CASES, build_case
from oracle import maxsim_oracle as O

RTOL = 2e-5   # fp32 vs fp32, different summation order only


def test_known_answer_vector(golden_dir):
    """BaseModel.test_score (reference BaseModel.py:70-75) → [[21., 41.]]."""
    sc = np.load(os.path.join(golden_dir, "score_cases.npz"))
    assert sc["kat_score"].tolist() == [[21.0, 41.0]]
    got = O.score_allpairs(sc["kat_Q"], sc["kat_D"], np.ones((1, 2)), np.ones((2, 2)))
    assert got.tolist() == [[21.0, 41.0]]


@pytest.mark.parametrize("name", ["ap_small", "ap_mid", "ap_views"])
def test_score_allpairs_matches_reference(golden_dir, name):
    sc = np.load(os.path.join(golden_dir, "score_cases.npz"))
    got = O.score_allpairs(sc[name + "_Q"].astype(np.float32), sc[name + "_D"].astype(np.float32),
                           sc[name + "_qmask"], sc[name + "_dmask"])
    np.testing.assert_allclose(got, sc[name + "_score"], rtol=RTOL, atol=1e-6)


def _check_topk(got_pids, got_scores, ref_pids, ref_scores, tol):
    """Same scores position by position; same pids except inside groups of (near-)tied scores."""
    got_pids, ref_pids = np.asarray(got_pids), np.asarray(ref_pids)
    got_scores, ref_scores = np.asarray(got_scores, dtype=np.float64), np.asarray(ref_scores, dtype=np.float64)
    assert got_pids.shape == ref_pids.shape
    scale = np.maximum(np.abs(ref_scores), 1.0)
    assert np.all(np.abs(got_scores - ref_scores) <= tol * scale), np.abs(got_scores - ref_scores).max()
    bad = np.nonzero(got_pids != ref_pids)[0]
    for i in bad:
        # position i may differ only if the reference holds got_pids[i] at a position whose score ties with i
        j = np.nonzero(ref_pids == got_pids[i])[0]
        assert j.size >= 1, f"pid {got_pids[i]} not in the reference list"
        assert np.abs(ref_scores[j] - ref_scores[i]).min() <= 2 * tol * scale[i], (i, got_pids[i], ref_pids[i])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_rank_forward_matches_reference(golden_dir, case):
    g = np.load(os.path.join(golden_dir, f"rank_{case['name']}.npz"))
    index, queries, cands = build_case(case)
    store = O.pad_store(index.emb)
    assert store.shape[0] == int(g["store_rows"][0])
    pf = O.doclens_pfxsum(index.doclens)
    assert pf[-4:].tolist() == g["pfxsum_tail"].tolist()
    strides = O.compute_strides(index.doclens)
    assert strides == g["strides"].tolist()
    for qi, (Q, pids) in enumerate(zip(queries, cands)):
        Qt = np.transpose(Q[None], (0, 2, 1))                      # [1, dim, q_len]
        for dname, depth in case["depths"]:
            p, s, all_s = O.rank_forward(store, index.doclens, pf, strides, Qt, pids, depth=depth,
                                         return_all_scores=True)
            _check_topk(p, s, g[f"q{qi}_{dname}_pids"], g[f"q{qi}_{dname}_scores"], RTOL)
        # the independent formulation the CUDA kernel implements (exact doclen + zero-floor flag)
        exact = O.maxsim_exact(store, index.doclens, pf, strides, Q, pids)
        np.testing.assert_allclose(exact, all_s, rtol=RTOL, atol=2e-6)
        if case.get("output_D"):
            p, D, M = O.rank_forward(store, index.doclens, pf, strides, Qt, pids, depth=case["output_D"],
                                     output_D_embedding=True)
            assert p == g[f"q{qi}_D_pids"].tolist()
            assert np.array_equal(D.astype(np.float16), g[f"q{qi}_D_rows"])   # bit-exact rows
            assert np.array_equal(M, g[f"q{qi}_D_mask"])


def test_zero_floor_is_visible_in_edge_case(golden_dir):
    """Without the a12' floor rule the oracle would NOT match the reference on short docs."""
    case = next(c for c in CASES if c["name"] == "edge")
    g = np.load(os.path.join(golden_dir, "rank_edge.npz"))
    index, queries, cands = build_case(case)
    store, pf = O.pad_store(index.emb), O.doclens_pfxsum(index.doclens)
    strides = O.compute_strides(index.doclens)
    floored = O.maxsim_exact(store, index.doclens, pf, strides, queries[0], cands[0], use_floor=True)
    raw = O.maxsim_exact(store, index.doclens, pf, strides, queries[0], cands[0], use_floor=False)
    assert np.abs(floored - raw).max() > 1.0
    ref_sorted = np.sort(g["q0_all_scores"])[::-1]
    np.testing.assert_allclose(np.sort(floored)[::-1], ref_sorted, rtol=RTOL, atol=5e-6)
    # docs whose doclen IS a stride keep their negative scores
    assert g["q0_all_scores"].min() < -1.0


def test_store_layout_roundtrip(tmp_path):
    """write_index (reference layout) → oracle.load_store: identical rows, 512 zero tail rows, pfxsum."""
    from colbert_b200 import synthetic
    idx = synthetic.make_index(5, 57, dim=128, lo=1, hi=20, num_parts=11)   # 11 parts: 10.pt sorts after 9.pt
    synthetic.write_index(idx, str(tmp_path))
    parts, paths, _ = O.get_parts(str(tmp_path))
    assert parts == list(range(11)) and paths[10].endswith("/10.pt")
    store, doclens = O.load_store(str(tmp_path), 128)
    assert np.array_equal(doclens, idx.doclens)
    assert store.shape[0] == idx.num_tokens + O.TAIL_PAD_ROWS
    assert np.array_equal(store[: idx.num_tokens], idx.emb)
    assert not store[idx.num_tokens:].any()


def test_percentile_semantics():
    """kthvalue(int(p*len/100)) is 1-indexed; fewer than 4 docs make k == 0 at p=25 (torch raises)."""
    d = np.array([5, 1, 9, 3, 7, 2, 8, 4], dtype=np.int64)
    assert O.percentile_kth(d, 25) == 2 and O.percentile_kth(d, 50) == 4 and O.percentile_kth(d, 75) == 7
    assert O.compute_strides(d) == [2, 4, 7, 9]
    with pytest.raises(IndexError):
        O.percentile_kth(np.array([3, 4, 5], dtype=np.int64), 25)


def test_merge_topk_total_order():
    s0, p0 = np.array([5.0, 3.0, 3.0], np.float32), np.array([10, 4, 9])
    s1, p1 = np.array([5.0, 4.0, -np.inf], np.float32), np.array([2, 7, -1])
    pid, sc = O.merge_topk([s0, s1], [p0, p1], 4)
    assert pid.tolist() == [2, 10, 7, 4] and sc.tolist() == [5.0, 5.0, 4.0, 3.0]


@pytest.mark.parametrize("case", CASES[:2], ids=[c["name"] for c in CASES[:2]])
def test_torch_port_matches_reference(golden_dir, case):
    """oracle/ref_port_torch.py (the timed CPU baseline) reproduces the reference's rankings."""
    import torch
    from oracle.ref_port_torch import CpuRankerPort
    g = np.load(os.path.join(golden_dir, f"rank_{case['name']}.npz"))
    index, queries, cands = build_case(case)
    port = CpuRankerPort(torch.from_numpy(O.pad_store(index.emb)), index.doclens.tolist(), max_candidates=256)
    assert port.strides == g["strides"].tolist()
    for qi, (Q, pids) in enumerate(zip(queries, cands)):
        Qt = torch.from_numpy(Q).unsqueeze(0).permute(0, 2, 1)
        p, s = port.rank_forward(Qt, pids.tolist(), depth=None)
        _check_topk(p, s, g[f"q{qi}_all_pids"], g[f"q{qi}_all_scores"], RTOL)


@pytest.mark.parametrize("name", ["g_small", "g_mid", "g_views", "g_wide", "g_tiles"])
def test_score_allpairs_grad_matches_reference_autograd(golden_dir, name):
    """oracle backward of BaseModel.score == the reference's own ops under torch autograd (BaseModel.py:41-45)."""
    g = np.load(os.path.join(golden_dir, "score_grad_cases.npz"))
    Q, D = g[f"{name}_Q"].astype(np.float32), g[f"{name}_D"].astype(np.float32)
    s, dQ, dD, _ = O.score_allpairs_grad(Q, D, g[f"{name}_qmask"], g[f"{name}_dmask"], g[f"{name}_W"])
    np.testing.assert_allclose(s, g[f"{name}_score"], rtol=0, atol=2e-5)
    # a different arg-max on a near-tie would show up as an O(1) difference: the tolerance only covers summation order
    np.testing.assert_allclose(dQ, g[f"{name}_dQ"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(dD, g[f"{name}_dD"], rtol=0, atol=2e-5)
