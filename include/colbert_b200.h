/*
 * colbert_b200.h — C ABI of the B200-native late-interaction scoring path.
 *
 * Drop-in boundary for ONE path of wuyaoxuehun/colbert: doclen-offset gather of per-token document
 * embeddings from the flat index store → MaxSim (Q·Dᵀ, max over document tokens / views, sum over
 * query tokens) → top-k.  The reference has no FFI for this path (it is pure Python calling torch);
 * the seam this library sits behind is the pair of Python call signatures
 *     ColbertRanker.rank_forward(Q, pids, views, depth, output_D_embedding)   colbert/ranking/colbert_ranker.py:75-137
 *     BaseModel.score(Q, D, q_mask, d_mask)                                   colbert/modeling/BaseModel.py:39-46
 * and every entry point below names the reference lines whose work it replaces.  INTEGRATION.md
 * shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++ / torch types; every function returns a cbk_status
 *     (0 = OK, negative = error) and never throws; cbk_last_error() gives the message of the last
 *     failure on the calling thread.
 *   - every pointer named d_* is DEVICE memory owned by the caller; nothing is allocated, freed or
 *     retained by the library; scratch comes from a caller-supplied workspace whose size is given by
 *     the matching *_workspace_bytes().
 *   - every launch takes an explicit cudaStream_t (passed as void*; 0 = legacy default stream) and is
 *     asynchronous with respect to the host.  Calls are thread-safe when they use distinct
 *     workspaces.
 *   - the library targets sm_100a only and reports CBK_ERR_UNSUPPORTED on any other device; there is
 *     no CPU fallback.
 *
 * Store layout (reference colbert_ranker.py:61-73, loaders.py:7-32): one flat row-major matrix
 * [n_store_rows, dim] of 16-bit floats (fp16 as the reference writes it, or bf16), the embeddings
 * of document p occupying rows pfxsum[p] .. pfxsum[p]+doclens[p]-1; the reference appends 512 zero
 * rows, this library does not need them but tolerates them (n_store_rows counts them).
 */
#ifndef COLBERT_B200_H
#define COLBERT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBK_ABI_VERSION 3
#define CBK_MAX_STRIDES 8      /* the reference produces at most 4 (percentiles 25/50/75 + max) */
#define CBK_MAX_QLEN 32        /* query rows per launch; longer queries are split by the caller  */
#define CBK_FLAG_BF16_NATIVE_MMA 1
#define CBK_FLAG_RERANK_TCGEN05 4       /* score with the tcgen05 / TMEM kernel instead of the mma.sync one */
#define CBK_FLAG_SKIP_FOREIGN_PIDS 2   /* sharded stores: a pid outside this shard scores -inf, not NaN */
#define CBK_FLAG_RERANK_GENERIC 8       /* dim != 128: score with the CUDA-core kernel even where the tensor-core one applies */
#define CBK_FLAG_FIXED_DOCLEN 16        /* every document has exactly strides[0] rows (multi-view index): offsets are pid * strides[0] */
#define CBK_FLAG_RERANK_KSPLIT 32        /* dim = 192 … 1024: score with the K-split mma.sync kernel instead of the tcgen05 streaming one */
#define CBK_TOPK_NEG_INF_IS_PADDING 1  /* top-k: candidates scored -inf are dropped (sharded rerank) */

typedef enum cbk_status {
  CBK_OK = 0,
  CBK_ERR_INVALID_ARG = -1,
  CBK_ERR_UNSUPPORTED = -2,   /* shape/dtype/device outside what the kernels implement */
  CBK_ERR_CUDA = -3,          /* a CUDA runtime/driver call failed; see cbk_last_error() */
  CBK_ERR_WORKSPACE = -4      /* workspace missing or too small */
} cbk_status;

typedef enum cbk_dtype {
  CBK_F16 = 0,                /* IEEE half — the reference's store dtype (encoder.py:175) */
  CBK_BF16 = 1,
  CBK_F32 = 2                 /* only as a source/destination of cbk_mask_cast_rows */
} cbk_dtype;

typedef enum cbk_mask_dtype {
  CBK_MASK_NONE = 0,          /* no mask: plain cast */
  CBK_MASK_U8 = 1,            /* bool / uint8 */
  CBK_MASK_I64 = 2,           /* torch.long, what the reference passes (colbert_ranker.py:112) */
  CBK_MASK_F32 = 3
} cbk_mask_dtype;

/* Message of the last error raised on this thread ("" if none). Never NULL. */
const char* cbk_last_error(void);

/* CBK_ABI_VERSION the library was built with. */
int cbk_abi_version(void);

/* 1 when device `device` is an sm_100 part the kernels can run on, else 0 (or a negative cbk_status). */
int cbk_device_supported(int device);

/* How many kernels this library has launched in this process since load (for bench accounting). */
uint64_t cbk_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * MaxSim rerank — replaces colbert_ranker.py:88-118 (pid→(doclen,offset) lookup, stride-bucket
 * gather, H2D, fp16→fp32 cast, length mask) and BaseModel.py:41-45 (mask-mul, einsum, max, sum),
 * for a BATCH of queries each with its own candidate list (CSR).
 *
 *   score[c] = Σ_{m<q_len} max( max_{t<doclen[p]} Q[q(c)][m]·store[pfxsum[p]+t] , floor[p] ),  p = cand_pids[c]
 *   floor[p] = 0 when doclens[p] is NOT one of `strides` (the reference's multiplicative mask gives a
 *              padded slot similarity 0, SURVEY.md §8 a12′), −inf otherwise; n_strides == 0 ⇒ no floor.
 *
 *   d_store        [n_store_rows, dim] store_dtype, 256-byte aligned base
 *   d_pfxsum       [n_docs + 1] int64    (colbert_ranker.py:32)
 *   d_doclens      [n_docs] int32
 *   pid_base       first GLOBAL pid held by this store (0 unless the corpus is sharded by pid range,
 *                  SURVEY.md §8e); candidates carry global pids, document p is local row block p - pid_base
 *   strides        host array of n_strides ints (colbert_ranker.py:36-40), may be NULL when n_strides == 0
 *   d_Q            [n_queries, q_len, dim] fp32 row-major (the reference's Q.permute(0,2,1), l.111)
 *   d_q_lens       NULL, or [n_queries] int32: query q has only q_lens[q] (<= q_len) real rows, the rest of its q_len
 *                  slots is padding and is read as zero whatever it holds (a zero row adds exactly 0 to every score).
 *                  This is what replaces the reference server's per-query keep_nonzero / qd_mask_to_realinput
 *                  (colbert/training/dense_server_client.py:44-46) when queries of different lengths share a batch.
 *   d_cand_pids    [n_cand_total] int64, concatenated candidate lists
 *   d_cand_rowptr  [n_queries + 1] int64, query q owns candidates rowptr[q] .. rowptr[q+1]-1;
 *                  the number of candidates actually scored is rowptr[n_queries], read on the device;
 *                  n_cand_total only has to be an upper bound of it (it sizes the launch)
 *   d_out_scores   [n_cand_total] fp32, written at the candidate's own position (the reference's
 *                  un-permute, colbert_ranker.py:120-122, is therefore not needed)
 *   d_workspace    ≥ cbk_maxsim_rerank_workspace_bytes() bytes
 *
 *   flags          bit-or of CBK_FLAG_*.  CBK_FLAG_SKIP_FOREIGN_PIDS: a pid outside
 *                  [pid_base, pid_base + n_docs) scores -inf (it belongs to another shard) instead of NaN.
 *                  CBK_FLAG_BF16_NATIVE_MMA: by default a bf16 store is converted to fp16 in
 *                  registers (exact for values in fp16's normal range — ColBERT embeddings are
 *                  L2-normalised, BaseModel.py:26) and multiplied with the query rounded to fp16
 *                  (11 significant bits; measured worst relative score error 2e-4).  With the flag the
 *                  bf16 values are multiplied directly with the query rounded to bf16 (8 bits; 1.2e-3,
 *                  outside the 1e-3 parity tolerance but safe for stores beyond fp16's range).
 *                  CBK_FLAG_FIXED_DOCLEN (with n_strides == 1): the caller guarantees that every document has
 *                  exactly strides[0] rows — the layout of an enable_multiview index, where a document is its
 *                  d_view view embeddings (BaseModel.py:21-27).  Document p then starts at row (p - pid_base) *
 *                  strides[0]; d_pfxsum / d_doclens are not read (dim == 128 and dim 192 … 1024 kernels; ignored by the others).
 *
 * Supported: 1 ≤ q_len ≤ CBK_MAX_QLEN, n_store_rows < 2^31; dim == 128 runs the TMA + tensor-core kernel the
 * benchmarks quote; a multiple of 64 from 192 to 1024 (the author's configuration: 768) a tcgen05 streaming kernel
 * (ragged documents, or — CBK_FLAG_FIXED_DOCLEN — the author's fixed-length multi-view index without metadata lookups;
 * CBK_FLAG_RERANK_KSPLIT selects the older K-split mma.sync kernel, which also serves dim 64); any other dim in [1, 1536] a generic
 * CUDA-core kernel with the same results contract (fp32 arithmetic; CBK_FLAG_RERANK_GENERIC forces it).  pids are range-checked on
 * the device; an out-of-range pid yields NaN at its position.
 * ------------------------------------------------------------------------------------------------ */
size_t cbk_maxsim_rerank_workspace_bytes(void);

int cbk_maxsim_rerank(const void* d_store, int store_dtype, int64_t n_store_rows, int dim,
                      const int64_t* d_pfxsum, const int32_t* d_doclens, int64_t n_docs, int64_t pid_base,
                      const int32_t* strides, int n_strides,
                      const float* d_Q, const int32_t* d_q_lens, int q_len, int64_t n_queries,
                      const int64_t* d_cand_pids, const int64_t* d_cand_rowptr, int64_t n_cand_total,
                      float* d_out_scores, void* d_workspace, size_t workspace_bytes, int flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-query top-k — replaces colbert_ranker.py:128-130 (full sort, descending, truncate to depth).
 * Total order: score descending, then pid ascending (the reference's torch.sort is unstable, so its
 * order among exact ties is unspecified).  A NaN score (cbk_maxsim_rerank's answer for a pid outside the
 * index) sorts LAST, below -inf, so an invalid candidate can never displace a real one.
 *
 *   d_scores, d_cand_pids, d_cand_rowptr as above; each query may hold at most
 *   cbk_topk_max_candidates() candidates; pids must be < 2^32.  flags: CBK_TOPK_NEG_INF_IS_PADDING.
 *   d_out_scores [n_queries, k] fp32, d_out_pids [n_queries, k] int64; when a query has fewer than k
 *   candidates the tail is filled with (−inf, −1).
 * ------------------------------------------------------------------------------------------------ */
int64_t cbk_topk_max_candidates(void);

int cbk_topk_per_query(const float* d_scores, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr,
                       int64_t n_queries, int64_t max_cand_per_query, int k, int flags,
                       float* d_out_scores, int64_t* d_out_pids, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One reference-shaped call with HOST buffers — the whole of ColbertRanker.rank_forward
 * (colbert_ranker.py:75-137) for one query: stage query + candidate pids through page-locked memory,
 * one host→device copy, cbk_maxsim_rerank, cbk_topk_per_query, one device→host copy of the k winners,
 * and a synchronise of `stream`; the function returns with h_out_* filled.
 *
 *   h_Q          fp32 query: [q_len, dim] row-major, or — q_dim_major != 0 — [dim, q_len], the layout
 *                the reference passes (Q[0] of its [1, dim, q_len] tensor, colbert_ranker.py:78)
 *   h_pids       [n] int64 candidates, 1 ≤ n ≤ cbk_topk_max_candidates(); k ≤ n
 *   h_out_pids   [k] int64, h_out_scores [k] fp32, score-descending (order of cbk_topk_per_query)
 *   d_scratch    device scratch and h_pinned page-locked host scratch, each of at least
 *                cbk_rank_forward_scratch_bytes(n, q_len, dim, k) bytes; owned by the caller, reusable
 *                across calls, not shared between concurrent calls
 *   store / metadata / strides / pid_base / flags as for cbk_maxsim_rerank
 * ------------------------------------------------------------------------------------------------ */
size_t cbk_rank_forward_scratch_bytes(int64_t n, int q_len, int dim, int k);

int cbk_rank_forward_host(const void* d_store, int store_dtype, int64_t n_store_rows, int dim,
                          const int64_t* d_pfxsum, const int32_t* d_doclens, int64_t n_docs, int64_t pid_base,
                          const int32_t* strides, int n_strides,
                          const float* h_Q, int q_len, int q_dim_major, const int64_t* h_pids, int64_t n, int k,
                          int64_t* h_out_pids, float* h_out_scores,
                          void* d_scratch, void* h_pinned, size_t scratch_bytes, int flags, void* stream);

/* Same selection, but the k winners of each query are returned as packed 64-bit keys
 * [ordered(score) : 32 | ~pid : 32] (0 = padding) — the unit the shards exchange: one all-gather of
 * [n_queries, k] uint64 per rank instead of separate score and pid buffers.  d_out_keys [n_queries, k]. */
int cbk_topk_per_query_keys(const float* d_scores, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr,
                            int64_t n_queries, int64_t max_cand_per_query, int k, int flags,
                            uint64_t* d_out_keys, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Merge of per-shard top-k lists (no counterpart in the reference, whose ranker is
 * single-GPU): d_keys [world, n_queries, k_in] packed keys as all-gathered from the ranks → the global
 * top-k per query under the same total order, decoded to d_out_scores / d_out_pids [n_queries, k]
 * ((−inf, −1) padding).  world * k_in ≤ cbk_topk_max_candidates().
 * ------------------------------------------------------------------------------------------------ */
int cbk_merge_topk_keys(const uint64_t* d_keys, int world, int64_t n_queries, int k_in, int k,
                        float* d_out_scores, int64_t* d_out_pids, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row gather — replaces colbert_ranker.py:105-109 as exposed by rank_forward(output_D_embedding=True)
 * (l.131-136): for each pid, `stride` consecutive store rows starting at pfxsum[pid] (rows past the
 * document's own length belong to the next document, exactly as the reference's stride-view reads
 * them), upcast to fp32, plus the length mask  mask[i][t] = (t + 1 <= doclens[pid_i]).
 *
 *   d_out_D [n, stride, dim] fp32, d_out_mask [n, stride] uint8 (0/1).  Rows at or past n_store_rows read as 0.
 * ------------------------------------------------------------------------------------------------ */
int cbk_gather_rows(const void* d_store, int store_dtype, int64_t n_store_rows, int dim,
                    const int64_t* d_pfxsum, const int32_t* d_doclens, int64_t n_docs,
                    const int64_t* d_pids, int64_t n, int stride,
                    float* d_out_D, uint8_t* d_out_mask, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Candidate routing for a sharded store: keep, per query and in order, the candidates
 * whose GLOBAL pid lies in [pid_lo, pid_hi).
 *   d_cand_pids [n_cand_total] int64, d_cand_rowptr [n_queries+1] int64   (replicated inputs)
 *   d_out_pids  [≥ n_cand_total] int64  (only the first d_out_rowptr[n_queries] entries are written)
 *   d_out_rowptr [n_queries+1] int64
 *   d_workspace ≥ cbk_partition_workspace_bytes(n_queries) bytes
 * The kept count stays on the device; cbk_maxsim_rerank reads it from d_cand_rowptr[n_queries], so the
 * outputs can be passed straight on (with n_cand_total as the upper bound) without a host sync.
 * ------------------------------------------------------------------------------------------------ */
size_t cbk_partition_workspace_bytes(int64_t n_queries);

int cbk_partition_candidates(const int64_t* d_cand_pids, const int64_t* d_cand_rowptr, int64_t n_queries,
                             int64_t pid_lo, int64_t pid_hi, int64_t* d_out_pids, int64_t* d_out_rowptr,
                             void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Candidate-generation post-processing — the step just before the scoring path:
 *   cbk_build_emb2pid           replaces ColbertIndex.build_emb2pid (colbert_ranker.py:163-174):
 *                               d_emb2pid[t] = document owning store row t, int32 [n_tokens].
 *   cbk_embedding_ids_to_pids   replaces ColbertIndex.embedding_ids_to_pids + uniq (colbert_ranker.py:212-235):
 *                               d_emb_ids [n_queries, n_ids] int64 (the ANN search's neighbour ids, -1 = none) →
 *                               per query the SORTED UNIQUE pids as CSR (d_out_pids capacity n_queries*n_ids,
 *                               d_out_rowptr [n_queries+1]); the reference returns the same set in Python
 *                               `set` order.  n_ids ≤ 16384 (= 32 query tokens × faiss_depth 512).  The CSR pair
 *                               feeds cbk_maxsim_rerank directly, with no host round trip.
 * ------------------------------------------------------------------------------------------------ */
int cbk_build_emb2pid(const int64_t* d_pfxsum, int64_t n_docs, int32_t* d_emb2pid, void* stream);
size_t cbk_embedding_ids_to_pids_workspace_bytes(int64_t n_queries, int n_ids);
int cbk_embedding_ids_to_pids(const int64_t* d_emb_ids, int64_t n_queries, int n_ids, const int32_t* d_emb2pid,
                              int64_t n_tokens, int64_t* d_out_pids, int64_t* d_out_rowptr, void* d_workspace,
                              size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked row cast — the multiplicative masks of BaseModel.score (BaseModel.py:41-42: D * d_mask[...,None],
 * Q * q_mask[...,None]) fused with the cast to the tensor-core input type:
 *     out[r, :] = (out_dtype)( (float)src[r, :] * (float)mask[r] )          mask == NULL ⇒ plain cast
 *   d_src [n_rows, dim] of src_dtype (CBK_F16 | CBK_BF16 | CBK_F32), d_mask [n_rows] of mask_dtype,
 *   d_out [n_rows, dim] of out_dtype.
 * ------------------------------------------------------------------------------------------------ */
int cbk_mask_cast_rows(const void* d_src, int src_dtype, int64_t n_rows, int dim, const void* d_mask, int mask_dtype,
                       void* d_out, int out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * All-pairs MaxSim of padded batches, forward and backward — BaseModel.score as the reference trains with it
 * (BaseModel.py:39-46 called from colbert/modeling/colbert_model.py:87-95 on the all-gathered Q / D of
 * colbert/training/training_utils.py:35-45), at any width that is a multiple of 64 up to 1024 (the author's: 768).
 *
 *   d_Qp  [n_queries, m, dim], d_Dp [n_docs, n, dim]: the MASKED operands in 16 bits (dtype CBK_F16 | CBK_BF16), i.e. the
 *         outputs of cbk_mask_cast_rows(Q, q_mask) and cbk_mask_cast_rows(D, d_mask) — BaseModel.py:41-42 fused with the
 *         cast.  A masked slot is a zero row and scores exactly 0, as in the reference.  m <= CBK_MAX_QLEN.
 *   cbk_score_allpairs_fwd   d_out_scores [n_queries, n_docs] fp32 = Σ_m max_n Qp[q,m]·Dp[d,n] (fp32 accumulation on
 *                            tcgen05 tensor cores; `simmat` of BaseModel.py:43 is never written);
 *                            d_out_argmax [n_queries, n_docs, m] int32 = the maximising n (first one on ties, like
 *                            torch.max on the CPU), or NULL when no backward pass will follow.
 *   cbk_score_allpairs_bwd   gradients of Σ_{q,d} grad_scores[q,d]·scores[q,d] with respect to the UNMASKED inputs:
 *                              d_grad_Q [n_queries, m, dim] fp32 = q_mask[q,m] · Σ_d grad[q,d] · Dp[d, argmax[q,d,m]]
 *                              d_grad_D [n_docs, n, dim]   fp32 = d_mask[d,n] · Σ_{(q,m): argmax[q,d,m]=n} grad[q,d] · Qp[q,m]
 *                            (what autograd derives for BaseModel.py:41-45); either may be NULL.  Masks as for
 *                            cbk_mask_cast_rows (CBK_MASK_NONE: all ones).  Sums run in a fixed order: results are
 *                            bit-reproducible.  d_grad_D needs n_queries * m <= 65535.
 * ------------------------------------------------------------------------------------------------ */
int cbk_score_allpairs_fwd(const void* d_Qp, const void* d_Dp, int dtype, int64_t n_queries, int m, int64_t n_docs, int n, int dim,
                           float* d_out_scores, int32_t* d_out_argmax, void* stream);
int cbk_score_allpairs_bwd(const void* d_Qp, const void* d_Dp, int dtype, int64_t n_queries, int m, int64_t n_docs, int n, int dim,
                           const float* d_grad_scores, const int32_t* d_argmax, const void* d_q_mask, int q_mask_dtype,
                           const void* d_d_mask, int d_mask_dtype, float* d_grad_Q, float* d_grad_D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exhaustive, query-batched MaxSim — every document of the store against a batch of queries: the
 * all-pairs shape of BaseModel.score (BaseModel.py:39-46) applied to the flat store of
 * colbert_ranker.py:61-73.  Document tiles are read from HBM once per pass of
 * up to 16 queries (8 for bf16 stores, whose query is split in hi + lo parts) and multiplied on tcgen05
 * tensor cores with TMEM accumulators.
 *
 *   cbk_build_doc_end_bits   index-time metadata: bit t of d_bits set ⇔ store row t is the last row of a
 *                            document; d_bits has cbk_doc_end_bits_bytes(n_store_rows) bytes.  Every
 *                            document must have at least one row.
 *   cbk_maxsim_exhaustive    d_out_scores [n_queries, n_docs] fp32 row-major: score of query q against
 *                            document p at [q*n_docs + p], same floor rule as cbk_maxsim_rerank.
 *                            d_Q [n_queries, q_len ≤ 32, 128] fp32; dim must be 128.
 *   cbk_topk_dense           per-row top-k of a dense score matrix [n_queries, n_docs] (pid = pid_base +
 *                            column), same total order as cbk_topk_per_query; k ≤ 8192.  as_keys != 0: the
 *                            winners are written to d_out_pids as packed keys (see cbk_topk_per_query_keys)
 *                            for the cross-shard merge, and d_out_scores may be NULL.
 * ------------------------------------------------------------------------------------------------ */
size_t cbk_doc_end_bits_bytes(int64_t n_store_rows);
int cbk_build_doc_end_bits(const int64_t* d_pfxsum, int64_t n_docs, int64_t n_store_rows, uint32_t* d_bits, void* stream);

size_t cbk_maxsim_exhaustive_workspace_bytes(int64_t n_queries);
int cbk_maxsim_exhaustive(const void* d_store, int store_dtype, int64_t n_store_rows, int dim,
                          const int64_t* d_pfxsum, const uint32_t* d_doc_end_bits, int64_t n_docs,
                          const int32_t* strides, int n_strides, const float* d_Q, int q_len, int64_t n_queries,
                          float* d_out_scores, void* d_workspace, size_t workspace_bytes, int flags, void* stream);

size_t cbk_topk_dense_workspace_bytes(int64_t n_queries, int64_t n_docs, int k);
int cbk_topk_dense(const float* d_scores, int64_t n_queries, int64_t n_docs, int k, int64_t pid_base, int as_keys,
                   float* d_out_scores, int64_t* d_out_pids, void* d_workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COLBERT_B200_H */
