"""Host-side mirror of the reference interface, CPU-checkable parts: index-directory readers,
percentile strides, and that the product package never imports the oracle."""
import ast
import os

import numpy as np
import pytest
import torch

from colbert_b200 import synthetic
from colbert_b200.indexing.index_manager import IndexManager, load_index_part
from colbert_b200.indexing.loaders import get_parts, load_doclens
from colbert_b200.ranking.colbert_ranker import torch_percentile
from oracle import maxsim_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_loaders_match_oracle(tmp_path):
    idx = synthetic.make_index(9, 40, dim=128, lo=1, hi=12, num_parts=12)
    synthetic.write_index(idx, str(tmp_path))
    assert get_parts(str(tmp_path)) == O.get_parts(str(tmp_path))
    assert load_doclens(str(tmp_path), flatten=True) == O.load_doclens(str(tmp_path), flatten=True)
    assert load_doclens(str(tmp_path), flatten=False) == O.load_doclens(str(tmp_path), flatten=False)
    part = load_index_part(os.path.join(str(tmp_path), "11.pt"))
    assert part.dtype == torch.float16 and part.shape[1] == 128


def test_get_parts_asserts_contiguous_numbering(tmp_path):
    torch.save(torch.zeros(1, 4), tmp_path / "0.pt")
    torch.save(torch.zeros(1, 4), tmp_path / "2.pt")
    with pytest.raises(AssertionError):
        get_parts(str(tmp_path))


def test_legacy_list_part_is_concatenated(tmp_path):
    a, b = torch.ones(2, 4, dtype=torch.float16), torch.zeros(3, 4, dtype=torch.float16)
    IndexManager(4).save([a, b], str(tmp_path / "0.pt"))
    assert load_index_part(str(tmp_path / "0.pt")).shape == (5, 4)


def test_torch_percentile_matches_oracle():
    rng = np.random.default_rng(0)
    for n in (4, 7, 100, 1001):
        d = rng.integers(1, 181, size=n)
        for p in (25, 50, 75, 100):
            assert torch_percentile(torch.from_numpy(d), p) == O.percentile_kth(d, p)


def test_product_never_imports_the_oracle():
    """colbert_b200/ and benchmarks/ must not import, call or link anything under oracle/ (it is test infrastructure;
    only tests/, __graft_entry__.smoke() and bench.py's CPU arm may)."""
    roots = [os.path.join(ROOT, "colbert_b200"), os.path.join(ROOT, "benchmarks")]
    for dirpath, _, files in (w for r in roots for w in os.walk(r)):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), (f, names)


def test_flat_store_roundtrip(tmp_path):
    """convert_index: reference layout → store.bin + doclens.i32 + meta.json; load_flat reads any pid range back
    bit-exactly in the reference's in-memory layout (512 zero tail rows)."""
    from colbert_b200.indexing.flat_store import convert_index, load_flat, read_meta
    idx = synthetic.make_index(12, 90, dim=128, lo=1, hi=30, num_parts=4)
    src, dst = tmp_path / "ref", tmp_path / "flat"
    synthetic.write_index(idx, str(src))
    meta = convert_index(str(src), str(dst))
    assert meta == {"dim": 128, "dtype": "float16", "num_docs": 90, "num_embeddings": idx.num_tokens, "parts": 4}
    m2, dl = read_meta(str(dst))
    assert m2 == meta and np.array_equal(dl, idx.doclens)
    assert os.path.getsize(dst / "store.bin") == idx.num_tokens * 256
    store, doclens, lo, _ = load_flat(str(dst), "cpu")
    assert np.array_equal(store.numpy(), O.pad_store(idx.emb)) and np.array_equal(doclens.numpy(), idx.doclens)
    pf = O.doclens_pfxsum(idx.doclens)
    store, doclens, lo, _ = load_flat(str(dst), "cpu", 17, 53)
    assert lo == 17 and np.array_equal(doclens.numpy(), idx.doclens[17:53])
    assert np.array_equal(store.numpy()[:-512], idx.emb[pf[17]: pf[53]]) and not store.numpy()[-512:].any()


def test_index_part_loader_validates_and_maps(tmp_path):
    """load_index_part / IndexManager: list parts, width and shape checks, memory-mapped load."""
    from colbert_b200.indexing.index_manager import IndexManager, load_index_part
    good = torch.randn(7, 16).half()
    IndexManager(16).save(good, str(tmp_path / "0.pt"))
    assert torch.equal(load_index_part(str(tmp_path / "0.pt"), verbose=False), good)
    assert torch.equal(load_index_part(str(tmp_path / "0.pt"), verbose=False, mmap=True, dim=16), good)
    with pytest.raises(ValueError):
        load_index_part(str(tmp_path / "0.pt"), verbose=False, dim=32)              # the ranker was built for another width
    with pytest.raises(ValueError):
        IndexManager(8).save(good, str(tmp_path / "1.pt"))                          # refuses to write a foreign width
    torch.save(torch.arange(10), str(tmp_path / "2.pt"))
    with pytest.raises(ValueError):
        load_index_part(str(tmp_path / "2.pt"), verbose=False)                      # not a [rows, dim] float matrix
    torch.save({"a": 1}, str(tmp_path / "3.pt"))
    with pytest.raises(TypeError):
        load_index_part(str(tmp_path / "3.pt"), verbose=False)
    torch.save([good[:3], good[3:]], str(tmp_path / "4.pt"))                        # legacy list-of-batches part
    assert torch.equal(load_index_part(str(tmp_path / "4.pt"), verbose=False), good)


def test_part_loader_does_not_fall_back_to_full_pickle(tmp_path, monkeypatch):
    """A part the weights-only unpickler refuses is NOT retried with pickle unless the caller opts in."""
    import pickle

    class Payload:                                        # stands in for "anything the safe loader refuses"
        def __reduce__(self):
            return (list, ([1.0],))

    path = str(tmp_path / "0.pt")
    with open(path, "wb") as fh:
        pickle.dump(Payload(), fh)
    monkeypatch.delenv("COLBERT_B200_ALLOW_PICKLE", raising=False)
    with pytest.raises(RuntimeError, match="weights-only"):
        load_index_part(path, verbose=False)
    # opting in reaches the unsafe loader (which then fails on the content checks, not on the load policy)
    with pytest.raises((TypeError, ValueError, RuntimeError)) as ei:
        load_index_part(path, verbose=False, allow_pickle=True)
    assert "weights-only" not in str(ei.value)


def test_get_representation_multiview_switch():
    """BaseModel.get_representation (reference BaseModel.py:21-27): with enable_multiview the first q_view / d_view
    hidden states are kept, projected and L2-normalised; without it every position is."""
    from types import SimpleNamespace
    from colbert_b200.modeling.BaseModel import BaseModel
    torch.manual_seed(0)
    m = BaseModel()
    m.linear = torch.nn.Linear(24, 16, bias=False)
    hidden = torch.randn(3, 40, 24)
    mv = SimpleNamespace(q_view=8, d_view=16)
    for enable in (True, False):
        m.args = SimpleNamespace(enable_multiview=enable, dense_multiview_args=mv)
        for is_query in (True, False):
            out = m.get_representation(hidden, is_query)
            rows = (mv.q_view if is_query else mv.d_view) if enable else 40
            expect = torch.nn.functional.normalize(hidden[:, :rows] @ m.linear.weight.t(), p=2, dim=2)
            assert tuple(out.shape) == (3, rows, 16)
            torch.testing.assert_close(out, expect)
            torch.testing.assert_close(out.norm(dim=2), torch.ones(3, rows))      # unit rows ⇒ every q·d ∈ [-1, 1]
    # against the reference's own method when the reference tree is here (this container; absent on the GPU box)
    ref_root = "/root/reference"
    if os.path.isdir(os.path.join(ref_root, "colbert")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_BaseModel", os.path.join(ref_root, "colbert/modeling/BaseModel.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        r = ref.BaseModel()
        r.linear = m.linear
        for enable in (True, False):
            r.args = m.args = SimpleNamespace(enable_multiview=enable, dense_multiview_args=mv)
            for is_query in (True, False):
                assert torch.equal(r.get_representation(hidden, is_query), m.get_representation(hidden, is_query))


def test_strides_setter_refreshes_the_ctypes_copy():
    """ColbertRanker.strides is a property: assigning the corpus-wide list (what ShardedColbertRanker does) rebuilds
    the ctypes array the single-call path passes to the library (a stale copy applied the wrong floor)."""
    from colbert_b200.ranking.colbert_ranker import ColbertRanker
    r = ColbertRanker.__new__(ColbertRanker)
    r.strides = [3, 9]
    assert list(r._strides_c) == [3, 9] and r.strides == [3, 9]
    r.strides = [5, 10, 20, 40]
    assert list(r._strides_c) == [5, 10, 20, 40] and r._views is None
