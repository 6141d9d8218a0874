"""The C-ABI library: builds in-tree for sm_100a, loads, and exports every symbol the header
declares.  No compute calls (no GPU here); argument validation that happens before any CUDA call is
exercised too."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "colbert_b200.h")
PROBE_HEADER = os.path.join(ROOT, "include", "colbert_b200_probe.h")


def declared_symbols(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cbk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("cbk_maxsim_rerank", "cbk_topk_per_query", "cbk_gather_rows", "cbk_mask_cast_rows",
                 "cbk_topk_per_query_keys", "cbk_merge_topk_keys", "cbk_partition_candidates", "cbk_maxsim_exhaustive", "cbk_build_emb2pid", "cbk_embedding_ids_to_pids", "cbk_build_doc_end_bits", "cbk_topk_dense", "cbk_last_error", "cbk_abi_version"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/colbert_b200.h but not exported"


def test_binding_covers_every_declared_symbol(built_lib):
    from colbert_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.cbk_abi_version() == 3
    assert lib.cbk_topk_max_candidates() == 16384          # the reference's BSIZE
    assert lib.cbk_maxsim_rerank_workspace_bytes() >= 4


def test_probe_library_is_separate(built_lib):
    """the tcgen05 self-test / rate probes are not in the product library; their own library exports what its header declares"""
    from colbert_b200 import _lib
    product = C.CDLL(built_lib)
    probe = _lib.load_probe()
    syms = declared_symbols(PROBE_HEADER)
    assert syms == sorted(_lib.PROBE_SIGNATURES) and syms
    for name in syms:
        assert hasattr(probe, name) and not hasattr(product, name)


def test_library_is_sm100a_and_uses_tma(built_lib):
    """The kernels are compiled for sm_100a and the rerank kernel really stages through TMA."""
    elf = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass          # cp.async.bulk.tensor
    assert "HMMA" in sass or "UTCHMMA" in sass


def test_invalid_arguments_are_reported_not_crashed(built_lib):
    from colbert_b200 import _lib
    lib = _lib.load()
    rc = lib.cbk_maxsim_rerank(None, 0, 10, 128, None, None, 1, 0, None, 0, None, None, 32, 1, None, None, 0, None, None, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.cbk_last_error()
    rc = lib.cbk_topk_per_query(None, None, None, 1, 1, 1, 0, None, None, None)
    assert rc == -1
    rc = lib.cbk_gather_rows(None, 0, 1, 128, None, None, 1, None, 1, 1, None, None, None)
    assert rc == -1
    rc = lib.cbk_rank_forward_host(None, 0, 10, 128, None, None, 1, 0, None, 0, None, 32, 1, None, 5, 3, None, None, None, None, 0, 0, None)
    assert rc == -1 and b"cbk_rank_forward_host" in lib.cbk_last_error()
    assert lib.cbk_rank_forward_scratch_bytes(1000, 32, 128, 10) >= 1000 * 12 + 32 * 128 * 4 + 256
    assert lib.cbk_rank_forward_scratch_bytes(0, 32, 128, 10) == 0
    rc = _lib.load_probe().cbk_selftest_umma_rate(128, 2, 100, 4, 1, None, None)   # no output buffer
    assert rc != 0
    with pytest.raises(_lib.CbkError):
        _lib.check("cbk_gather_rows", rc)


def test_product_fails_loudly_without_the_library(monkeypatch, tmp_path):
    from colbert_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_ranker_refuses_cpu_device():
    import torch
    from colbert_b200.ranking import ColbertRanker
    with pytest.raises(RuntimeError, match="no CPU path"):
        ColbertRanker.from_tensors(torch.zeros(4, 128, dtype=torch.float16), [4], device="cpu")
