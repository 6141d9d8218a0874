"""Thin tensor-level wrappers over the C ABI (include/colbert_b200.h).

Each function checks devices / dtypes / contiguity, passes raw device pointers, sizes and the
current CUDA stream to libcolbert_b200.so, and returns CUDA tensors.  Nothing here computes on the
host and nothing falls back to torch ops: a missing library or an unsupported shape raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need(t: torch.Tensor, name: str, dtype, device) -> None:
    if not t.is_cuda or t.device != device:
        raise ValueError(f"{name} must live on {device}, got {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


_workspaces = {}


def _workspace(device: torch.device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = int(_lib.load().cbk_maxsim_rerank_workspace_bytes())
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def maxsim_rerank(store: torch.Tensor, pfxsum: torch.Tensor, doclens: torch.Tensor, strides: Sequence[int],
                  Q: torch.Tensor, cand_pids: torch.Tensor, cand_rowptr: torch.Tensor,
                  out: Optional[torch.Tensor] = None, flags: int = 0, pid_base: int = 0,
                  q_lens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scores[c] for the CSR candidate lists; see cbk_maxsim_rerank in include/colbert_b200.h.

    store [rows, dim] fp16|bf16 · pfxsum [n_docs+1] int64 · doclens [n_docs] int32 · Q [B, q_len, dim] fp32
    cand_pids [n] int64 · cand_rowptr [B+1] int64 · q_lens None | [B] int32 (real rows per query; the rest of the
    q_len slots is padding)  → fp32 [n] (all on the store's device)."""
    lib = _lib.load()
    dev = store.device
    _need(store, "store", store.dtype, dev)
    _need(pfxsum, "pfxsum", torch.int64, dev)
    _need(doclens, "doclens", torch.int32, dev)
    _need(Q, "Q", torch.float32, dev)
    _need(cand_pids, "cand_pids", torch.int64, dev)
    _need(cand_rowptr, "cand_rowptr", torch.int64, dev)
    if Q.dim() != 3 or store.dim() != 2 or Q.size(2) != store.size(1):
        raise ValueError(f"Q must be [B, q_len, dim={store.size(1)}], got {tuple(Q.shape)}")
    n_q, q_len, dim = Q.shape
    if cand_rowptr.numel() != n_q + 1:
        raise ValueError("cand_rowptr must have B+1 entries")
    if q_lens is not None:
        _need(q_lens, "q_lens", torch.int32, dev)
        if q_lens.numel() != n_q:
            raise ValueError("q_lens must have one entry per query")
    n = cand_pids.numel()
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=dev)
    else:
        _need(out, "out", torch.float32, dev)
        if out.numel() != n:
            raise ValueError("out must have one entry per candidate")
    st = (C.c_int32 * max(1, len(strides)))(*[int(s) for s in strides])
    ws = _workspace(dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_maxsim_rerank(_ptr(store), _lib.dtype_code(store.dtype), store.size(0), dim, _ptr(pfxsum),
                                   _ptr(doclens), doclens.numel(), int(pid_base), C.cast(st, C.c_void_p),
                                   len(strides), _ptr(Q), _ptr(q_lens),
                                   q_len, n_q, _ptr(cand_pids), _ptr(cand_rowptr), n, _ptr(out), _ptr(ws),
                                   ws.numel(), int(flags), C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_maxsim_rerank", rc)
    return out


def topk_per_query(scores: torch.Tensor, cand_pids: torch.Tensor, cand_rowptr: torch.Tensor, k: int,
                   max_cand_per_query: int, flags: int = 0, as_keys: bool = False):
    """(top scores [B,k] fp32, top pids [B,k] int64), score-descending, ties by ascending pid;
    lists shorter than k are padded with (-inf, -1).  With ``as_keys`` the winners come back as packed
    int64 keys [B,k] (see cbk_topk_per_query_keys) ready for an all-gather."""
    lib = _lib.load()
    dev = scores.device
    _need(scores, "scores", torch.float32, dev)
    _need(cand_pids, "cand_pids", torch.int64, dev)
    _need(cand_rowptr, "cand_rowptr", torch.int64, dev)
    n_q = cand_rowptr.numel() - 1
    stream = C.c_void_p(_lib.current_stream_ptr(dev))
    with torch.cuda.device(dev):
        if as_keys:
            out_k = torch.empty((n_q, k), dtype=torch.int64, device=dev)
            rc = lib.cbk_topk_per_query_keys(_ptr(scores), _ptr(cand_pids), _ptr(cand_rowptr), n_q,
                                             int(max_cand_per_query), int(k), int(flags), _ptr(out_k), stream)
            _lib.check("cbk_topk_per_query_keys", rc)
            return out_k
        out_s = torch.empty((n_q, k), dtype=torch.float32, device=dev)
        out_p = torch.empty((n_q, k), dtype=torch.int64, device=dev)
        rc = lib.cbk_topk_per_query(_ptr(scores), _ptr(cand_pids), _ptr(cand_rowptr), n_q, int(max_cand_per_query),
                                    int(k), int(flags), _ptr(out_s), _ptr(out_p), stream)
    _lib.check("cbk_topk_per_query", rc)
    return out_s, out_p


def merge_topk_keys(keys: torch.Tensor, k: int):
    """keys [W, B, k_in] int64 (packed, rank-major, as all-gathered) → (scores [B,k] fp32, pids [B,k] int64):
    the global top-k per query under the total order (score desc, pid asc).  See cbk_merge_topk_keys."""
    lib = _lib.load()
    dev = keys.device
    _need(keys, "keys", torch.int64, dev)
    W, B, k_in = keys.shape
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_p = torch.empty((B, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_merge_topk_keys(_ptr(keys), W, B, k_in, int(k), _ptr(out_s), _ptr(out_p),
                                     C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_merge_topk_keys", rc)
    return out_s, out_p


def gather_rows(store: torch.Tensor, pfxsum: torch.Tensor, doclens: torch.Tensor, pids: torch.Tensor, stride: int):
    """(D fp32 [n, stride, dim], mask bool [n, stride]) — bit-exact copies of `stride` store rows per
    pid starting at its offset (reference colbert_ranker.py:105-109).  See cbk_gather_rows."""
    lib = _lib.load()
    dev = store.device
    _need(store, "store", store.dtype, dev)
    _need(pfxsum, "pfxsum", torch.int64, dev)
    _need(doclens, "doclens", torch.int32, dev)
    _need(pids, "pids", torch.int64, dev)
    n, dim = pids.numel(), store.size(1)
    D = torch.empty((n, stride, dim), dtype=torch.float32, device=dev)
    mask = torch.empty((n, stride), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_gather_rows(_ptr(store), _lib.dtype_code(store.dtype), store.size(0), dim, _ptr(pfxsum),
                                 _ptr(doclens), doclens.numel(), _ptr(pids), n, int(stride), _ptr(D), _ptr(mask),
                                 C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_gather_rows", rc)
    return D, mask.view(torch.bool)


def mask_cast_rows(src: torch.Tensor, mask: Optional[torch.Tensor], out_dtype: torch.dtype) -> torch.Tensor:
    """out[r, :] = out_dtype(float(src[r, :]) * float(mask[r])) for a 2-D ``src`` (fp16/bf16/fp32) and a
    1-D per-row mask (bool/uint8/int64/fp32, or None for a plain cast).  See cbk_mask_cast_rows."""
    lib = _lib.load()
    dev = src.device
    _need(src, "src", src.dtype, dev)
    if src.dim() != 2:
        raise ValueError("src must be 2-D [rows, dim]")
    n_rows, dim = src.shape
    mcode = _lib.CBK_MASK_NONE
    if mask is not None:
        _need(mask, "mask", mask.dtype, dev)
        if mask.numel() != n_rows:
            raise ValueError("mask must have one entry per row")
        mcode = _lib.mask_code(mask.dtype)
    out = torch.empty((n_rows, dim), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_mask_cast_rows(_ptr(src), _lib.dtype_code(src.dtype, True), n_rows, dim, _ptr(mask), mcode,
                                    _ptr(out), _lib.dtype_code(out_dtype, True),
                                    C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_mask_cast_rows", rc)
    return out


def score_allpairs_supported(m: int, dim: int) -> bool:
    """Shapes cbk_score_allpairs_fwd / _bwd implement: at most CBK_MAX_QLEN query rows, width a multiple of 64 up to 1024."""
    return 0 < m <= _lib.CBK_MAX_QLEN and dim % 64 == 0 and 0 < dim <= 1024


def score_allpairs_fwd(Qp: torch.Tensor, Dp: torch.Tensor, want_argmax: bool = True):
    """scores [q, d] fp32 (and argmax [q, d, m] int32) of the masked 16-bit operands Qp [q, m, h], Dp [d, n, h];
    see cbk_score_allpairs_fwd."""
    lib = _lib.load()
    dev = Dp.device
    _need(Dp, "Dp", Dp.dtype, dev)
    _need(Qp, "Qp", Dp.dtype, dev)
    nq, m, h = Qp.shape
    nd, n, h2 = Dp.shape
    if h != h2:
        raise ValueError(f"widths differ: {tuple(Qp.shape)} vs {tuple(Dp.shape)}")
    scores = torch.empty((nq, nd), dtype=torch.float32, device=dev)
    argmax = torch.empty((nq, nd, m), dtype=torch.int32, device=dev) if want_argmax else None
    with torch.cuda.device(dev):
        rc = lib.cbk_score_allpairs_fwd(_ptr(Qp), _ptr(Dp), _lib.dtype_code(Dp.dtype), nq, m, nd, n, h, _ptr(scores),
                                        _ptr(argmax), C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_score_allpairs_fwd", rc)
    return scores, argmax


def score_allpairs_bwd(Qp: torch.Tensor, Dp: torch.Tensor, grad_scores: torch.Tensor, argmax: torch.Tensor,
                       q_mask: Optional[torch.Tensor], d_mask: Optional[torch.Tensor], want_dq: bool = True,
                       want_dd: bool = True):
    """(grad_Q [q, m, h], grad_D [d, n, h]) fp32 for the upstream gradient grad_scores [q, d]; see cbk_score_allpairs_bwd."""
    lib = _lib.load()
    dev = Dp.device
    _need(Dp, "Dp", Dp.dtype, dev)
    _need(Qp, "Qp", Dp.dtype, dev)
    _need(grad_scores, "grad_scores", torch.float32, dev)
    _need(argmax, "argmax", torch.int32, dev)
    nq, m, h = Qp.shape
    nd, n, _ = Dp.shape
    if tuple(grad_scores.shape) != (nq, nd) or tuple(argmax.shape) != (nq, nd, m):
        raise ValueError("grad_scores must be [q, d] and argmax [q, d, m]")
    codes = []
    for name, mk, rows in (("q_mask", q_mask, nq * m), ("d_mask", d_mask, nd * n)):
        if mk is None:
            codes.append(_lib.CBK_MASK_NONE)
            continue
        _need(mk, name, mk.dtype, dev)
        if mk.numel() != rows:
            raise ValueError(f"{name} must have one entry per row")
        codes.append(_lib.mask_code(mk.dtype))
    dq = torch.empty((nq, m, h), dtype=torch.float32, device=dev) if want_dq else None
    dd = torch.empty((nd, n, h), dtype=torch.float32, device=dev) if want_dd else None
    with torch.cuda.device(dev):
        rc = lib.cbk_score_allpairs_bwd(_ptr(Qp), _ptr(Dp), _lib.dtype_code(Dp.dtype), nq, m, nd, n, h, _ptr(grad_scores),
                                        _ptr(argmax), _ptr(q_mask), codes[0], _ptr(d_mask), codes[1], _ptr(dq), _ptr(dd),
                                        C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_score_allpairs_bwd", rc)
    return dq, dd


def partition_candidates(cand_pids: torch.Tensor, cand_rowptr: torch.Tensor, pid_lo: int, pid_hi: int):
    """Keep, per query and in order, the candidates with pid in [pid_lo, pid_hi) → (pids [n_total] int64 of
    which the first rowptr[-1] are valid, rowptr [B+1] int64); no host synchronisation.
    See cbk_partition_candidates."""
    lib = _lib.load()
    dev = cand_pids.device
    _need(cand_pids, "cand_pids", torch.int64, dev)
    _need(cand_rowptr, "cand_rowptr", torch.int64, dev)
    n_q = cand_rowptr.numel() - 1
    out_p = torch.empty_like(cand_pids)
    out_r = torch.empty_like(cand_rowptr)
    ws = torch.empty(int(lib.cbk_partition_workspace_bytes(n_q)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_partition_candidates(_ptr(cand_pids), _ptr(cand_rowptr), n_q, int(pid_lo), int(pid_hi), _ptr(out_p),
                                          _ptr(out_r), _ptr(ws), ws.numel(), C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_partition_candidates", rc)
    return out_p, out_r


def selftest_umma_gemm(A: torch.Tensor, B: torch.Tensor, tma_3d: bool = False) -> torch.Tensor:
    """C[128, N] = A[128,128] · B[N,128]^T through TMA → tcgen05.mma → TMEM → tcgen05.ld (one CTA).
    ``tma_3d``: load each operand with ONE 3-D TMA op instead of one op per 64-column half."""
    lib = _lib.load_probe()
    dev = A.device
    _need(A, "A", A.dtype, dev)
    _need(B, "B", B.dtype, dev)
    assert A.shape == (128, 128) and B.dim() == 2 and B.size(1) == 128
    N = B.size(0)
    Cout = torch.empty((128, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_selftest_umma_gemm(_ptr(A), _ptr(B), N, int(A.dtype == torch.bfloat16) | (2 if tma_3d else 0),
                                        int(B.dtype == torch.bfloat16),
                                        _ptr(Cout), C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_selftest_umma_gemm", rc)
    return Cout


def selftest_umma_rate(N: int, mode: int, iters: int, n_acc: int, ctas_per_sm: int = 1, device=None) -> torch.Tensor:
    """Cycles each CTA took to issue and retire ``iters`` tiles of [128, N] += A[128,128] · B[N,128]^T
    (cbk_selftest_umma_rate; mode 0 = both operands from shared memory, 1 = A from TMEM)."""
    lib = _lib.load_probe()
    dev = torch.device(device if device is not None else "cuda:0")
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    out = torch.zeros(n_sm * ctas_per_sm, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_selftest_umma_rate(int(N), int(mode), int(iters), int(n_acc), int(ctas_per_sm), _ptr(out),
                                        C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_selftest_umma_rate", rc)
    return out


def build_doc_end_bits(pfxsum: torch.Tensor, n_store_rows: int) -> torch.Tensor:
    """Index-time metadata for the exhaustive kernel: one bit per store row, set on the last row of each
    document (int32 words).  See cbk_build_doc_end_bits."""
    lib = _lib.load()
    dev = pfxsum.device
    _need(pfxsum, "pfxsum", torch.int64, dev)
    bits = torch.empty(int(lib.cbk_doc_end_bits_bytes(int(n_store_rows))) // 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_build_doc_end_bits(_ptr(pfxsum), pfxsum.numel() - 1, int(n_store_rows), _ptr(bits),
                                        C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_build_doc_end_bits", rc)
    return bits


def maxsim_exhaustive(store: torch.Tensor, pfxsum: torch.Tensor, doc_end_bits: torch.Tensor, strides: Sequence[int],
                      Q: torch.Tensor, flags: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scores [B, n_docs] fp32 of every document against every query (tcgen05 kernel); see cbk_maxsim_exhaustive."""
    lib = _lib.load()
    dev = store.device
    _need(store, "store", store.dtype, dev)
    _need(pfxsum, "pfxsum", torch.int64, dev)
    _need(doc_end_bits, "doc_end_bits", torch.int32, dev)
    _need(Q, "Q", torch.float32, dev)
    n_q, q_len, dim = Q.shape
    n_docs = pfxsum.numel() - 1
    if out is None:
        out = torch.empty((n_q, n_docs), dtype=torch.float32, device=dev)
    ws = torch.empty(int(lib.cbk_maxsim_exhaustive_workspace_bytes(n_q)), dtype=torch.uint8, device=dev)
    st = (C.c_int32 * max(1, len(strides)))(*[int(s) for s in strides])
    with torch.cuda.device(dev):
        rc = lib.cbk_maxsim_exhaustive(_ptr(store), _lib.dtype_code(store.dtype), store.size(0), dim, _ptr(pfxsum),
                                       _ptr(doc_end_bits), n_docs, C.cast(st, C.c_void_p), len(strides), _ptr(Q), q_len,
                                       n_q, _ptr(out), _ptr(ws), ws.numel(), int(flags),
                                       C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_maxsim_exhaustive", rc)
    return out


def topk_dense(scores: torch.Tensor, k: int, pid_base: int = 0, as_keys: bool = False):
    """Row-wise top-k of a dense [B, n_docs] score matrix → (scores [B,k], pids [B,k]); pid = pid_base + column.
    With ``as_keys`` → packed int64 keys [B,k] for the cross-shard merge."""
    lib = _lib.load()
    dev = scores.device
    _need(scores, "scores", torch.float32, dev)
    n_q, n_docs = scores.shape
    ws = torch.empty(int(lib.cbk_topk_dense_workspace_bytes(n_q, n_docs, int(k))), dtype=torch.uint8, device=dev)
    out_s = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    out_p = torch.empty((n_q, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_topk_dense(_ptr(scores), n_q, n_docs, int(k), int(pid_base), int(as_keys),
                                None if as_keys else _ptr(out_s), _ptr(out_p), _ptr(ws), ws.numel(),
                                C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_topk_dense", rc)
    return out_p if as_keys else (out_s, out_p)


def build_emb2pid(pfxsum: torch.Tensor) -> torch.Tensor:
    """int32 [n_tokens]: the document owning each store row (reference ColbertIndex.build_emb2pid)."""
    lib = _lib.load()
    dev = pfxsum.device
    _need(pfxsum, "pfxsum", torch.int64, dev)
    n_tokens = int(pfxsum[-1].item())
    out = torch.empty(n_tokens, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_build_emb2pid(_ptr(pfxsum), pfxsum.numel() - 1, _ptr(out), C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_build_emb2pid", rc)
    return out


def embedding_ids_to_pids(emb_ids: torch.Tensor, emb2pid: torch.Tensor):
    """emb_ids [B, n_ids] int64 (-1 = no neighbour) → per query the sorted unique pids as CSR
    (pids [B*n_ids] int64 of which rowptr[-1] are valid, rowptr [B+1] int64).  See cbk_embedding_ids_to_pids."""
    lib = _lib.load()
    dev = emb_ids.device
    _need(emb_ids, "emb_ids", torch.int64, dev)
    _need(emb2pid, "emb2pid", torch.int32, dev)
    B, n_ids = emb_ids.shape
    out_p = torch.empty(B * n_ids, dtype=torch.int64, device=dev)
    out_r = torch.empty(B + 1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.cbk_embedding_ids_to_pids_workspace_bytes(B, n_ids)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cbk_embedding_ids_to_pids(_ptr(emb_ids), B, n_ids, _ptr(emb2pid), emb2pid.numel(), _ptr(out_p), _ptr(out_r),
                                           _ptr(ws), ws.numel(), C.c_void_p(_lib.current_stream_ptr(dev)))
    _lib.check("cbk_embedding_ids_to_pids", rc)
    return out_p, out_r
