"""ctypes binding of libcolbert_b200.so — the C ABI declared in include/colbert_b200.h.

There is NO fallback: if the shared object has not been built (``python -m colbert_b200.csrc.build``
or ``__graft_entry__.build()``) importing the symbols raises, and every wrapper raises
:class:`CbkError` when the library reports a non-zero status.  torch is used only to obtain device
pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# COLBERT_B200_LIB: another build of the same library (A/B experiments on kernel variants); default = the in-tree build
LIB_PATH = os.environ.get("COLBERT_B200_LIB") or os.path.join(_HERE, "csrc", "libcolbert_b200.so")

CBK_F16, CBK_BF16, CBK_F32 = 0, 1, 2
CBK_MASK_NONE, CBK_MASK_U8, CBK_MASK_I64, CBK_MASK_F32 = 0, 1, 2, 3
CBK_MAX_STRIDES = 8
CBK_MAX_QLEN = 32
CBK_FLAG_BF16_NATIVE_MMA = 1
CBK_FLAG_SKIP_FOREIGN_PIDS = 2
CBK_FLAG_RERANK_TCGEN05 = 4
CBK_FLAG_RERANK_GENERIC = 8
CBK_FLAG_FIXED_DOCLEN = 16
CBK_FLAG_RERANK_KSPLIT = 32
CBK_ABI_VERSION = 3
CBK_TOPK_NEG_INF_IS_PADDING = 1

# name → (restype, argtypes); mirrors include/colbert_b200.h one to one
_vp, _i64, _i32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_size_t
SIGNATURES = {
    "cbk_last_error": (C.c_char_p, []),
    "cbk_abi_version": (C.c_int, []),
    "cbk_device_supported": (C.c_int, [C.c_int]),
    "cbk_launch_count": (C.c_uint64, []),
    "cbk_maxsim_rerank_workspace_bytes": (_sz, []),
    "cbk_maxsim_rerank": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp, _i32, _i64,
                                    _vp, _vp, _i64, _vp, _vp, _sz, _i32, _vp]),
    "cbk_topk_max_candidates": (_i64, []),
    "cbk_topk_per_query": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    "cbk_rank_forward_scratch_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "cbk_rank_forward_host": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _i32, _vp, _i64,
                                        _i32, _vp, _vp, _vp, _vp, _sz, _i32, _vp]),
    "cbk_topk_per_query_keys": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "cbk_merge_topk_keys": (C.c_int, [_vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    "cbk_gather_rows": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    "cbk_partition_workspace_bytes": (_sz, [_i64]),
    "cbk_partition_candidates": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "cbk_build_emb2pid": (C.c_int, [_vp, _i64, _vp, _vp]),
    "cbk_embedding_ids_to_pids_workspace_bytes": (_sz, [_i64, _i32]),
    "cbk_embedding_ids_to_pids": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "cbk_doc_end_bits_bytes": (_sz, [_i64]),
    "cbk_build_doc_end_bits": (C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "cbk_maxsim_exhaustive_workspace_bytes": (_sz, [_i64]),
    "cbk_maxsim_exhaustive": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _i32, _vp, _i32, _i64, _vp, _vp, _sz,
                                        _i32, _vp]),
    "cbk_topk_dense_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "cbk_topk_dense": (C.c_int, [_vp, _i64, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "cbk_mask_cast_rows": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _i32, _vp, _i32, _vp]),
    "cbk_score_allpairs_fwd": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    "cbk_score_allpairs_bwd": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _i32,
                                         _vp, _vp, _vp]),
}


# include/colbert_b200_probe.h — libcolbert_b200_probe.so (tests / benchmarks only, not the product library)
PROBE_LIB_PATH = os.path.join(_HERE, "csrc", "libcolbert_b200_probe.so")
PROBE_SIGNATURES = {
    "cbk_selftest_umma_gemm": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "cbk_selftest_umma_rate": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _vp, _vp]),
}


class CbkError(RuntimeError):
    """A call into libcolbert_b200.so returned a non-zero status."""

    def __init__(self, fn: str, status: int, message: str):
        super().__init__(f"{fn} failed with status {status}: {message}")
        self.fn, self.status, self.message = fn, status, message


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load (once) and return the shared library; raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -m colbert_b200.csrc.build` (nvcc, sm_100a). colbert_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)       # the probe library resolves its cbk:: helpers against this one
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.cbk_abi_version() != CBK_ABI_VERSION:
        raise RuntimeError(f"ABI version mismatch: library reports {lib.cbk_abi_version()}, binding expects {CBK_ABI_VERSION}")
    _lib = lib
    return lib


_probe: Optional[C.CDLL] = None


def load_probe() -> C.CDLL:
    """The probe library (self-test and issue-rate probes of the tcgen05 building blocks); the product library is loaded
    first — the probes report errors and count launches through it."""
    global _probe
    if _probe is not None:
        return _probe
    load()
    if not os.path.exists(PROBE_LIB_PATH):
        raise RuntimeError(f"{PROBE_LIB_PATH} is missing: run `python -m colbert_b200.csrc.build`")
    lib = C.CDLL(PROBE_LIB_PATH)
    for name, (res, args) in PROBE_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _probe = lib
    return lib


def check(fn: str, status: int) -> None:
    if status != 0:
        raise CbkError(fn, status, load().cbk_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().cbk_launch_count())


def dtype_code(torch_dtype, allow_f32: bool = False) -> int:
    import torch
    if torch_dtype == torch.float16:
        return CBK_F16
    if torch_dtype == torch.bfloat16:
        return CBK_BF16
    if allow_f32 and torch_dtype == torch.float32:
        return CBK_F32
    raise TypeError(f"dtype must be float16 or bfloat16{' or float32' if allow_f32 else ''}, got {torch_dtype}")


def mask_code(torch_dtype) -> int:
    import torch
    if torch_dtype in (torch.bool, torch.uint8):
        return CBK_MASK_U8
    if torch_dtype == torch.int64:
        return CBK_MASK_I64
    if torch_dtype == torch.float32:
        return CBK_MASK_F32
    raise TypeError(f"mask dtype must be bool, uint8, int64 or float32, got {torch_dtype}")


def current_stream_ptr(device) -> int:
    import torch
    return int(torch.cuda.current_stream(device).cuda_stream)
