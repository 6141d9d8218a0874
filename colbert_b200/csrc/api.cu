// extern "C" entry points of libcolbert_b200.so (declared in include/colbert_b200.h) plus the small
// amount of host state they share: thread-local last-error string, launch counter, cached SM count,
// and the driver entry point used to encode TMA tensor maps (resolved at run time so that the
// library links against the CUDA runtime only).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "cbk_common.cuh"

namespace cbk {

// dispatchers implemented in the kernel translation units
int rerank_dispatch(const void*, int, int64_t, int, const int64_t*, const int32_t*, int64_t, int64_t, const int32_t*, int,
                    const float*, const int32_t*, int, int64_t, const int64_t*, const int64_t*, int64_t, float*, void*, int, cudaStream_t);
int topk_dispatch(const float*, const int64_t*, const int64_t*, int64_t, int64_t, int, int, float*, int64_t*, uint64_t*,
                  cudaStream_t);
int merge_dispatch(const uint64_t*, int, int64_t, int, int, float*, int64_t*, cudaStream_t);
int64_t topk_max_candidates();
int gather_dispatch(const void*, int, int64_t, int, const int64_t*, const int32_t*, int64_t, const int64_t*, int64_t,
                    int, float*, uint8_t*, cudaStream_t);

int partition_dispatch(const int64_t*, const int64_t*, int64_t, int64_t, int64_t, int64_t*, int64_t*, void*, cudaStream_t);
size_t partition_workspace_bytes(int64_t);
size_t doc_end_bits_bytes(int64_t);
int doc_end_bits_dispatch(const int64_t*, int64_t, int64_t, uint32_t*, cudaStream_t);
size_t exhaustive_workspace_bytes(int64_t);
int exhaustive_dispatch(const void*, int, int64_t, const int64_t*, const uint32_t*, int64_t, const int32_t*, int, const float*,
                        int, int64_t, float*, void*, int, cudaStream_t);
size_t topk_dense_workspace_bytes(int64_t, int64_t, int);
int topk_dense_dispatch(const float*, int64_t, int64_t, int, int64_t, int, float*, int64_t*, void*, cudaStream_t);
int emb2pid_dispatch(const int64_t*, int64_t, int32_t*, cudaStream_t);
int unique_pids_dispatch(const int64_t*, int64_t, int, const int32_t*, int64_t, int64_t*, int64_t*, void*, cudaStream_t);
int rerank_generic_dispatch(const void*, int, int64_t, int, const int64_t*, const int32_t*, int64_t, int64_t, const int32_t*, int,
                            const float*, const int32_t*, int, int64_t, const int64_t*, const int64_t*, float*, int, cudaStream_t);
bool rerank_wide_supports(int dim);
int rerank_wide_dispatch(const void*, int, int64_t, int, const int64_t*, const int32_t*, int64_t, int64_t, const int32_t*, int,
                         const float*, const int32_t*, int, int64_t, const int64_t*, const int64_t*, int64_t, float*, void*, int, cudaStream_t);
bool rerank_wide_stream_supports(int, int, int);
int rerank_wide_stream_dispatch(const void*, int, int64_t, int, const int64_t*, const int32_t*, int64_t, int64_t, const int32_t*, int,
                                const float*, const int32_t*, int, int64_t, const int64_t*, const int64_t*, int64_t, float*, int,
                                cudaStream_t);
int rerank_umma_dispatch(const void*, int, int64_t, int, const int64_t*, const int32_t*, int64_t, int64_t, const int32_t*, int,
                         const float*, const int32_t*, int, int64_t, const int64_t*, const int64_t*, int64_t, float*, void*, int, cudaStream_t);
int mask_cast_dispatch(const void*, int, int64_t, int, const void*, int, void*, int, cudaStream_t);
int score_allpairs_fwd_dispatch(const void*, const void*, int, int64_t, int, int64_t, int, int, float*, int32_t*, cudaStream_t);
bool score_allpairs_bwd_fits(int64_t, int, int);
int score_allpairs_bwd_dispatch(const void*, const void*, int, int64_t, int, int64_t, int, int, const float*, const int32_t*, const void*,
                                int, const void*, int, float*, float*, cudaStream_t);

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || p == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_store_tensor_map(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the installed driver");
    return CBK_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim) * 2};  // bytes between consecutive rows
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estride[2] = {1, 1};
  // 16-bit payload: the TMA only moves bits, so fp16 and bf16 share one descriptor type
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p rows %lld dim %d box %dx%d)", static_cast<int>(r),
              base, static_cast<long long>(rows), dim, box_cols, box_rows);
    return CBK_ERR_CUDA;
  }
  return CBK_OK;
}

// One TMA op for a whole [rows, 128] tile in the layout the kernels use ([half][row][64 columns], 128-B swizzle):
// the store is described as 3-D {64 columns, rows, 2 halves} with byte strides {256 (row), 128 (half)} — the third
// dimension has the SMALLER stride — and a box of {64, box_rows, 2}.
int make_store_tensor_map_3d(CUtensorMap* out, const void* base, int64_t rows, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the installed driver");
    return CBK_ERR_CUDA;
  }
  const cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(rows), 2};
  const cuuint64_t gstride[2] = {256, 128};
  const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 2};
  const cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (rows %lld box_rows %d)", static_cast<int>(r),
              static_cast<long long>(rows), box_rows);
    return CBK_ERR_CUDA;
  }
  return CBK_OK;
}

// The same idea for any width that is a multiple of 64: {64 columns, rows, dim/64 slabs}, byte strides {2*dim (row),
// 128 (slab)}, box {64, box_rows, dim/64} ⇒ shared memory [slab][row][64 columns], each slab swizzled on its own.
int make_store_tensor_map_3d_wide(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the installed driver");
    return CBK_ERR_CUDA;
  }
  const cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(dim / 64)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(dim) * 2, 128};
  const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(dim / 64)};
  const cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D, dim %d) failed with CUresult %d (rows %lld box_rows %d)", dim, static_cast<int>(r),
              static_cast<long long>(rows), box_rows);
    return CBK_ERR_CUDA;
  }
  return CBK_OK;
}

int make_rows_tensor_map(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_rows) {
  return make_store_tensor_map(out, base, rows, dim, 64, box_rows);
}

// Packed queries [n_queries, m, dim] seen as {dim, m, n_queries} with a box of {64, rpq, 128 / rpq}: one op brings a K slab
// of a block of 4 queries x 32 row slots (or 8 x 16) as 128 rows ([query][row][64 columns], 128-B swizzle); rows >= m and
// queries >= n_queries are outside the tensor and arrive as zeros, so nothing has to be padded in memory.
int make_query_block_tensor_map(CUtensorMap* out, const void* base, int64_t n_queries, int m, int dim, int rows_per_query) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the installed driver");
    return CBK_ERR_CUDA;
  }
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(m), static_cast<cuuint64_t>(n_queries)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(dim) * 2, static_cast<cuuint64_t>(m) * dim * 2};
  const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(rows_per_query), static_cast<cuuint32_t>(128 / rows_per_query)};
  const cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (query blocks) failed with CUresult %d (queries %lld m %d dim %d)", static_cast<int>(r),
              static_cast<long long>(n_queries), m, dim);
    return CBK_ERR_CUDA;
  }
  return CBK_OK;
}

static int check_device() {
  int dev = 0;
  CBK_CUDA(cudaGetDevice(&dev));
  int major = 0;
  CBK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CBK_CHECK_SUPPORTED(major == 10, "device %d has compute capability major %d; this library is sm_100a only", dev, major);
  return CBK_OK;
}

}  // namespace cbk

using namespace cbk;

extern "C" {

const char* cbk_last_error(void) { return g_err; }

int cbk_abi_version(void) { return CBK_ABI_VERSION; }

int cbk_device_supported(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    return CBK_ERR_CUDA;
  }
  return major == 10 ? 1 : 0;
}

uint64_t cbk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t cbk_maxsim_rerank_workspace_bytes(void) { return 256; }

int cbk_maxsim_rerank(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                      const int32_t* d_doclens, int64_t n_docs, int64_t pid_base, const int32_t* strides, int n_strides,
                      const float* d_Q, const int32_t* d_q_lens,
                      int q_len, int64_t n_queries, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr,
                      int64_t n_cand_total, float* d_out_scores, void* d_workspace, size_t workspace_bytes,
                      int flags, void* stream) {
  const bool fixed_len = (flags & CBK_FLAG_FIXED_DOCLEN) && n_strides == 1 &&
                         (dim == 128 || rerank_wide_stream_supports(dim, q_len, flags));   // metadata arrays unused
  CBK_CHECK_ARG(d_store && (fixed_len || (d_pfxsum && d_doclens)) && d_Q && d_cand_pids && d_cand_rowptr && d_out_scores,
                "cbk_maxsim_rerank: null pointer argument");
  CBK_CHECK_ARG(store_dtype == CBK_F16 || store_dtype == CBK_BF16, "cbk_maxsim_rerank: unknown store dtype %d", store_dtype);
  CBK_CHECK_ARG(n_store_rows > 0 && n_docs > 0 && n_queries > 0 && n_cand_total >= 0,
                "cbk_maxsim_rerank: sizes must be positive (rows %lld docs %lld queries %lld cands %lld)",
                (long long)n_store_rows, (long long)n_docs, (long long)n_queries, (long long)n_cand_total);
  CBK_CHECK_ARG(n_strides >= 0 && n_strides <= CBK_MAX_STRIDES && (n_strides == 0 || strides),
                "cbk_maxsim_rerank: n_strides %d outside [0, %d] or strides is NULL", n_strides, CBK_MAX_STRIDES);
  CBK_CHECK_SUPPORTED(dim >= 1 && dim <= 1536, "cbk_maxsim_rerank: dim %d outside [1, 1536]", dim);
  CBK_CHECK_SUPPORTED(q_len >= 1 && q_len <= CBK_MAX_QLEN, "cbk_maxsim_rerank: q_len %d outside [1, %d]", q_len,
                      CBK_MAX_QLEN);
  CBK_CHECK_SUPPORTED(n_store_rows < (1ll << 31), "cbk_maxsim_rerank: store of %lld rows exceeds 2^31-1",
                      (long long)n_store_rows);
  CBK_CHECK_ARG(dim != 128 || (reinterpret_cast<uintptr_t>(d_store) & 0xff) == 0,
                "cbk_maxsim_rerank: store base must be 256-byte aligned");
  if (!d_workspace || workspace_bytes < cbk_maxsim_rerank_workspace_bytes()) {
    set_error("cbk_maxsim_rerank: workspace of %zu bytes, need %zu", workspace_bytes, cbk_maxsim_rerank_workspace_bytes());
    return CBK_ERR_WORKSPACE;
  }
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  if (n_cand_total == 0) return CBK_OK;
  if (rerank_wide_stream_supports(dim, q_len, flags) && (reinterpret_cast<uintptr_t>(d_store) & 0xf) == 0)
    return rerank_wide_stream_dispatch(d_store, store_dtype, n_store_rows, dim, d_pfxsum, d_doclens, n_docs, pid_base, strides,
                                       n_strides, d_Q, d_q_lens, q_len, n_queries, d_cand_pids, d_cand_rowptr, n_cand_total,
                                       d_out_scores, flags, static_cast<cudaStream_t>(stream));   // dim 256 ... 1024: tcgen05 streaming kernel
  if (dim != 128 && !(flags & CBK_FLAG_RERANK_GENERIC) && rerank_wide_supports(dim) &&
      (reinterpret_cast<uintptr_t>(d_store) & 0xf) == 0)
    return rerank_wide_dispatch(d_store, store_dtype, n_store_rows, dim, d_pfxsum, d_doclens, n_docs, pid_base, strides, n_strides,
                                d_Q, d_q_lens, q_len, n_queries, d_cand_pids, d_cand_rowptr, n_cand_total, d_out_scores, d_workspace,
                                flags, static_cast<cudaStream_t>(stream));   // K-split tensor-core kernel: dim = 64, 192, …, 768, 1024
  if (dim != 128)   // any other width (CUDA cores)
    return rerank_generic_dispatch(d_store, store_dtype, n_store_rows, dim, d_pfxsum, d_doclens, n_docs, pid_base, strides,
                                   n_strides, d_Q, d_q_lens, q_len, n_queries, d_cand_pids, d_cand_rowptr, d_out_scores, flags,
                                   static_cast<cudaStream_t>(stream));
  if (flags & CBK_FLAG_RERANK_TCGEN05)
    return rerank_umma_dispatch(d_store, store_dtype, n_store_rows, dim, d_pfxsum, d_doclens, n_docs, pid_base, strides,
                                n_strides, d_Q, d_q_lens, q_len, n_queries, d_cand_pids, d_cand_rowptr, n_cand_total, d_out_scores,
                                d_workspace, flags, static_cast<cudaStream_t>(stream));
  return rerank_dispatch(d_store, store_dtype, n_store_rows, dim, d_pfxsum, d_doclens, n_docs, pid_base, strides, n_strides, d_Q,
                         d_q_lens, q_len, n_queries, d_cand_pids, d_cand_rowptr, n_cand_total, d_out_scores, d_workspace, flags,
                         static_cast<cudaStream_t>(stream));
}

int64_t cbk_topk_max_candidates(void) { return topk_max_candidates(); }

static int topk_common(const char* fn, const float* d_scores, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr,
                       int64_t n_queries, int64_t max_cand_per_query, int k, int flags, float* d_out_scores,
                       int64_t* d_out_pids, uint64_t* d_out_keys, void* stream) {
  CBK_CHECK_ARG(d_scores && d_cand_pids && d_cand_rowptr && ((d_out_scores && d_out_pids) || d_out_keys),
                "%s: null pointer argument", fn);
  CBK_CHECK_ARG(n_queries > 0 && k > 0 && max_cand_per_query > 0, "%s: n_queries %lld, k %d, max_cand %lld", fn,
                (long long)n_queries, k, (long long)max_cand_per_query);
  CBK_CHECK_SUPPORTED(max_cand_per_query <= topk_max_candidates(), "%s: %lld candidates per query exceeds the limit of %lld",
                      fn, (long long)max_cand_per_query, (long long)topk_max_candidates());
  CBK_CHECK_SUPPORTED(n_queries < (1ll << 31), "%s: too many queries", fn);
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return topk_dispatch(d_scores, d_cand_pids, d_cand_rowptr, n_queries, max_cand_per_query, k, flags, d_out_scores,
                       d_out_pids, d_out_keys, static_cast<cudaStream_t>(stream));
}

int cbk_topk_per_query(const float* d_scores, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr,
                       int64_t n_queries, int64_t max_cand_per_query, int k, int flags, float* d_out_scores,
                       int64_t* d_out_pids, void* stream) {
  CBK_CHECK_ARG(d_out_scores && d_out_pids, "cbk_topk_per_query: null pointer argument");
  return topk_common("cbk_topk_per_query", d_scores, d_cand_pids, d_cand_rowptr, n_queries, max_cand_per_query, k, flags,
                     d_out_scores, d_out_pids, nullptr, stream);
}

int cbk_topk_per_query_keys(const float* d_scores, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr,
                            int64_t n_queries, int64_t max_cand_per_query, int k, int flags, uint64_t* d_out_keys,
                            void* stream) {
  CBK_CHECK_ARG(d_out_keys, "cbk_topk_per_query_keys: null pointer argument");
  return topk_common("cbk_topk_per_query_keys", d_scores, d_cand_pids, d_cand_rowptr, n_queries, max_cand_per_query, k,
                     flags, nullptr, nullptr, d_out_keys, stream);
}

// scratch layout shared by the device and the page-locked copy (offsets in bytes):
//   [0, 256) rerank workspace | rowptr 2 x i64 | Q q_len*dim f32 | pids n x i64 | out_pids k x i64 | out_scores k x f32 |
//   scores n x f32 (device only)
namespace {
struct RankForwardLayout {
  size_t rowptr, q, pids, out_pids, out_scores, scores, total;
};
RankForwardLayout rank_forward_layout(int64_t n, int q_len, int dim, int k) {
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  RankForwardLayout L;
  L.rowptr = 256;
  L.q = L.rowptr + 16;
  L.pids = up(L.q + sizeof(float) * q_len * dim);
  L.out_pids = up(L.pids + sizeof(int64_t) * n);
  L.out_scores = L.out_pids + sizeof(int64_t) * k;
  L.scores = up(L.out_scores + sizeof(float) * k);
  L.total = up(L.scores + sizeof(float) * n);
  return L;
}
}  // namespace

size_t cbk_rank_forward_scratch_bytes(int64_t n, int q_len, int dim, int k) {
  if (n < 1 || q_len < 1 || dim < 1 || k < 1) return 0;
  return rank_forward_layout(n, q_len, dim, k).total;
}

int cbk_rank_forward_host(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                          const int32_t* d_doclens, int64_t n_docs, int64_t pid_base, const int32_t* strides, int n_strides,
                          const float* h_Q, int q_len, int q_dim_major, const int64_t* h_pids, int64_t n, int k,
                          int64_t* h_out_pids, float* h_out_scores, void* d_scratch, void* h_pinned, size_t scratch_bytes,
                          int flags, void* stream) {
  CBK_CHECK_ARG(h_Q && h_pids && h_out_pids && h_out_scores && d_scratch && h_pinned,
                "cbk_rank_forward_host: null pointer argument");
  CBK_CHECK_ARG(n >= 1 && k >= 1 && k <= n, "cbk_rank_forward_host: need 1 <= k <= n (n %lld, k %d)", (long long)n, k);
  CBK_CHECK_SUPPORTED(q_len >= 1 && q_len <= CBK_MAX_QLEN && dim >= 1 && dim <= 1536,
                      "cbk_rank_forward_host: q_len %d outside [1, %d] or dim %d outside [1, 1536]", q_len, CBK_MAX_QLEN, dim);
  CBK_CHECK_SUPPORTED(n <= topk_max_candidates(), "cbk_rank_forward_host: %lld candidates exceed the limit of %lld",
                      (long long)n, (long long)topk_max_candidates());
  const RankForwardLayout L = rank_forward_layout(n, q_len, dim, k);
  if (scratch_bytes < L.total) {
    set_error("cbk_rank_forward_host: scratch of %zu bytes, need %zu", scratch_bytes, L.total);
    return CBK_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* hp = static_cast<char*>(h_pinned);
  char* dp = static_cast<char*>(d_scratch);
  // stage: rowptr, query (transposed to [q_len, dim] when it arrives dim-major), pids
  int64_t* h_rowptr = reinterpret_cast<int64_t*>(hp + L.rowptr);
  h_rowptr[0] = 0;
  h_rowptr[1] = n;
  float* hq = reinterpret_cast<float*>(hp + L.q);
  if (q_dim_major) {
    for (int d = 0; d < dim; ++d)
      for (int m = 0; m < q_len; ++m) hq[static_cast<size_t>(m) * dim + d] = h_Q[static_cast<size_t>(d) * q_len + m];
  } else {
    memcpy(hq, h_Q, sizeof(float) * q_len * dim);
  }
  memcpy(hp + L.pids, h_pids, sizeof(int64_t) * n);
  const size_t in_bytes = L.pids + sizeof(int64_t) * n - L.rowptr;
  CBK_CUDA(cudaMemcpyAsync(dp + L.rowptr, hp + L.rowptr, in_bytes, cudaMemcpyHostToDevice, st));
  const int64_t* d_rowptr = reinterpret_cast<const int64_t*>(dp + L.rowptr);
  const int64_t* d_pids = reinterpret_cast<const int64_t*>(dp + L.pids);
  float* d_scores = reinterpret_cast<float*>(dp + L.scores);
  int rc = cbk_maxsim_rerank(d_store, store_dtype, n_store_rows, dim, d_pfxsum, d_doclens, n_docs, pid_base, strides, n_strides,
                             reinterpret_cast<const float*>(dp + L.q), nullptr, q_len, 1, d_pids, d_rowptr, n, d_scores, dp, 256,
                             flags, stream);
  if (rc != CBK_OK) return rc;
  // the k winners go straight into the page-locked scratch when the device can address it (unified addressing maps
  // every cudaHostAlloc'ed buffer): one copy engine round trip less per call
  void* mapped = nullptr;
  const bool zero_copy = cudaHostGetDevicePointer(&mapped, hp, 0) == cudaSuccess && mapped != nullptr;
  if (!zero_copy) (void)cudaGetLastError();
  char* op = zero_copy ? static_cast<char*>(mapped) : dp;
  rc = cbk_topk_per_query(d_scores, d_pids, d_rowptr, 1, n, k, 0, reinterpret_cast<float*>(op + L.out_scores),
                          reinterpret_cast<int64_t*>(op + L.out_pids), stream);
  if (rc != CBK_OK) return rc;
  if (!zero_copy) {
    const size_t out_bytes = sizeof(int64_t) * k + sizeof(float) * k;
    CBK_CUDA(cudaMemcpyAsync(hp + L.out_pids, dp + L.out_pids, out_bytes, cudaMemcpyDeviceToHost, st));
  }
  CBK_CUDA(cudaStreamSynchronize(st));
  memcpy(h_out_pids, hp + L.out_pids, sizeof(int64_t) * k);
  memcpy(h_out_scores, hp + L.out_scores, sizeof(float) * k);
  return CBK_OK;
}

int cbk_merge_topk_keys(const uint64_t* d_keys, int world, int64_t n_queries, int k_in, int k, float* d_out_scores,
                        int64_t* d_out_pids, void* stream) {
  CBK_CHECK_ARG(d_keys && d_out_scores && d_out_pids, "cbk_merge_topk_keys: null pointer argument");
  CBK_CHECK_ARG(world > 0 && n_queries > 0 && k_in > 0 && k > 0, "cbk_merge_topk_keys: sizes must be positive");
  CBK_CHECK_SUPPORTED(static_cast<int64_t>(world) * k_in <= topk_max_candidates(),
                      "cbk_merge_topk_keys: world*k_in = %lld exceeds the limit of %lld", (long long)world * k_in,
                      (long long)topk_max_candidates());
  CBK_CHECK_SUPPORTED(n_queries < (1ll << 31), "cbk_merge_topk_keys: too many queries");
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return merge_dispatch(d_keys, world, n_queries, k_in, k, d_out_scores, d_out_pids, static_cast<cudaStream_t>(stream));
}

int cbk_gather_rows(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                    const int32_t* d_doclens, int64_t n_docs, const int64_t* d_pids, int64_t n, int stride,
                    float* d_out_D, uint8_t* d_out_mask, void* stream) {
  CBK_CHECK_ARG(d_store && d_pfxsum && d_doclens && d_pids && d_out_D && d_out_mask, "cbk_gather_rows: null pointer argument");
  CBK_CHECK_ARG(store_dtype == CBK_F16 || store_dtype == CBK_BF16, "cbk_gather_rows: unknown store dtype %d", store_dtype);
  CBK_CHECK_ARG(n > 0 && stride > 0 && n_docs > 0 && n_store_rows > 0, "cbk_gather_rows: sizes must be positive");
  CBK_CHECK_SUPPORTED(dim > 0 && dim % 8 == 0, "cbk_gather_rows: dim %d must be a multiple of 8", dim);
  CBK_CHECK_ARG((reinterpret_cast<uintptr_t>(d_store) & 0xf) == 0, "cbk_gather_rows: store base must be 16-byte aligned");
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return gather_dispatch(d_store, store_dtype, n_store_rows, dim, d_pfxsum, d_doclens, n_docs, d_pids, n, stride,
                         d_out_D, d_out_mask, static_cast<cudaStream_t>(stream));
}

size_t cbk_partition_workspace_bytes(int64_t n_queries) { return cbk::partition_workspace_bytes(n_queries); }

int cbk_partition_candidates(const int64_t* d_cand_pids, const int64_t* d_cand_rowptr, int64_t n_queries, int64_t pid_lo,
                             int64_t pid_hi, int64_t* d_out_pids, int64_t* d_out_rowptr, void* d_workspace,
                             size_t workspace_bytes, void* stream) {
  CBK_CHECK_ARG(d_cand_pids && d_cand_rowptr && d_out_pids && d_out_rowptr, "cbk_partition_candidates: null pointer argument");
  CBK_CHECK_ARG(n_queries > 0 && pid_lo <= pid_hi, "cbk_partition_candidates: n_queries %lld, range [%lld, %lld)",
                (long long)n_queries, (long long)pid_lo, (long long)pid_hi);
  if (!d_workspace || workspace_bytes < cbk_partition_workspace_bytes(n_queries)) {
    set_error("cbk_partition_candidates: workspace of %zu bytes, need %zu", workspace_bytes,
              cbk_partition_workspace_bytes(n_queries));
    return CBK_ERR_WORKSPACE;
  }
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return partition_dispatch(d_cand_pids, d_cand_rowptr, n_queries, pid_lo, pid_hi, d_out_pids, d_out_rowptr,
                            static_cast<int64_t*>(d_workspace), static_cast<cudaStream_t>(stream));
}

int cbk_mask_cast_rows(const void* d_src, int src_dtype, int64_t n_rows, int dim, const void* d_mask, int mask_dtype,
                       void* d_out, int out_dtype, void* stream) {
  CBK_CHECK_ARG(d_src && d_out, "cbk_mask_cast_rows: null pointer argument");
  CBK_CHECK_ARG(n_rows > 0 && dim > 0, "cbk_mask_cast_rows: sizes must be positive");
  CBK_CHECK_ARG(mask_dtype == CBK_MASK_NONE || d_mask, "cbk_mask_cast_rows: mask dtype %d given but mask is NULL", mask_dtype);
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return mask_cast_dispatch(d_src, src_dtype, n_rows, dim, mask_dtype == CBK_MASK_NONE ? nullptr : d_mask, mask_dtype,
                            d_out, out_dtype, static_cast<cudaStream_t>(stream));
}

static int check_allpairs_shape(const char* fn, int dtype, int64_t nq, int m, int64_t nd, int n, int dim) {
  CBK_CHECK_ARG(nq > 0 && m > 0 && nd > 0 && n > 0 && dim > 0, "%s: sizes must be positive", fn);
  CBK_CHECK_ARG(dtype == CBK_F16 || dtype == CBK_BF16, "%s: operands must be CBK_F16 or CBK_BF16 (got %d)", fn, dtype);
  CBK_CHECK_SUPPORTED(m <= CBK_MAX_QLEN, "%s: %d query rows; at most %d per call", fn, m, CBK_MAX_QLEN);
  CBK_CHECK_SUPPORTED(dim % 64 == 0 && dim <= 1024, "%s: dim %d; the tensor-core path needs a multiple of 64 up to 1024", fn, dim);
  CBK_CHECK_SUPPORTED(nd * n < (1ll << 31) - 512 && nq * m < (1ll << 31), "%s: more than 2^31 rows", fn);
  return CBK_OK;
}

int cbk_score_allpairs_fwd(const void* d_Qp, const void* d_Dp, int dtype, int64_t n_queries, int m, int64_t n_docs, int n, int dim,
                           float* d_out_scores, int32_t* d_out_argmax, void* stream) {
  CBK_CHECK_ARG(d_Qp && d_Dp && d_out_scores, "cbk_score_allpairs_fwd: null pointer argument");
  int rc = check_allpairs_shape("cbk_score_allpairs_fwd", dtype, n_queries, m, n_docs, n, dim);
  if (rc != CBK_OK) return rc;
  CBK_CHECK_ARG((reinterpret_cast<uintptr_t>(d_Qp) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_Dp) & 15) == 0,
                "cbk_score_allpairs_fwd: operands must be 16-byte aligned");
  rc = check_device();
  if (rc != CBK_OK) return rc;
  return score_allpairs_fwd_dispatch(d_Qp, d_Dp, dtype, n_queries, m, n_docs, n, dim, d_out_scores, d_out_argmax,
                                     static_cast<cudaStream_t>(stream));
}

int cbk_score_allpairs_bwd(const void* d_Qp, const void* d_Dp, int dtype, int64_t n_queries, int m, int64_t n_docs, int n, int dim,
                           const float* d_grad_scores, const int32_t* d_argmax, const void* d_q_mask, int q_mask_dtype,
                           const void* d_d_mask, int d_mask_dtype, float* d_grad_Q, float* d_grad_D, void* stream) {
  CBK_CHECK_ARG(d_Qp && d_Dp && d_grad_scores && d_argmax, "cbk_score_allpairs_bwd: null pointer argument");
  CBK_CHECK_ARG(d_grad_Q || d_grad_D, "cbk_score_allpairs_bwd: neither gradient requested");
  CBK_CHECK_ARG((q_mask_dtype == CBK_MASK_NONE || d_q_mask) && (d_mask_dtype == CBK_MASK_NONE || d_d_mask),
                "cbk_score_allpairs_bwd: mask dtype given but mask is NULL");
  int rc = check_allpairs_shape("cbk_score_allpairs_bwd", dtype, n_queries, m, n_docs, n, dim);
  if (rc != CBK_OK) return rc;
  CBK_CHECK_SUPPORTED(!d_grad_D || score_allpairs_bwd_fits(n_queries, m, n),
                      "cbk_score_allpairs_bwd: %lld query rows x %d document rows do not fit the per-document bucket sort "
                      "(at most 65535 query rows in total)", static_cast<long long>(n_queries * m), n);
  rc = check_device();
  if (rc != CBK_OK) return rc;
  return score_allpairs_bwd_dispatch(d_Qp, d_Dp, dtype, n_queries, m, n_docs, n, dim, d_grad_scores, d_argmax,
                                     q_mask_dtype == CBK_MASK_NONE ? nullptr : d_q_mask, q_mask_dtype,
                                     d_mask_dtype == CBK_MASK_NONE ? nullptr : d_d_mask, d_mask_dtype, d_grad_Q, d_grad_D,
                                     static_cast<cudaStream_t>(stream));
}

int cbk_build_emb2pid(const int64_t* d_pfxsum, int64_t n_docs, int32_t* d_emb2pid, void* stream) {
  CBK_CHECK_ARG(d_pfxsum && d_emb2pid && n_docs > 0, "cbk_build_emb2pid: bad argument");
  CBK_CHECK_SUPPORTED(n_docs < (1ll << 31), "cbk_build_emb2pid: too many documents for int32 pids");
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return emb2pid_dispatch(d_pfxsum, n_docs, d_emb2pid, static_cast<cudaStream_t>(stream));
}

size_t cbk_embedding_ids_to_pids_workspace_bytes(int64_t n_queries, int n_ids) {
  if (n_queries <= 0 || n_ids <= 0) return 8;
  return static_cast<size_t>(n_queries) * (static_cast<size_t>(n_ids) + 1) * sizeof(int64_t);
}

int cbk_embedding_ids_to_pids(const int64_t* d_emb_ids, int64_t n_queries, int n_ids, const int32_t* d_emb2pid,
                              int64_t n_tokens, int64_t* d_out_pids, int64_t* d_out_rowptr, void* d_workspace,
                              size_t workspace_bytes, void* stream) {
  CBK_CHECK_ARG(d_emb_ids && d_emb2pid && d_out_pids && d_out_rowptr, "cbk_embedding_ids_to_pids: null pointer argument");
  CBK_CHECK_ARG(n_queries > 0 && n_ids > 0 && n_tokens > 0, "cbk_embedding_ids_to_pids: sizes must be positive");
  CBK_CHECK_SUPPORTED(n_ids <= 16384 && n_queries < (1ll << 31), "cbk_embedding_ids_to_pids: n_ids %d exceeds 16384", n_ids);
  if (!d_workspace || workspace_bytes < cbk_embedding_ids_to_pids_workspace_bytes(n_queries, n_ids)) {
    set_error("cbk_embedding_ids_to_pids: workspace of %zu bytes, need %zu", workspace_bytes,
              cbk_embedding_ids_to_pids_workspace_bytes(n_queries, n_ids));
    return CBK_ERR_WORKSPACE;
  }
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return unique_pids_dispatch(d_emb_ids, n_queries, n_ids, d_emb2pid, n_tokens, d_out_pids, d_out_rowptr, d_workspace,
                              static_cast<cudaStream_t>(stream));
}

size_t cbk_doc_end_bits_bytes(int64_t n_store_rows) { return doc_end_bits_bytes(n_store_rows > 0 ? n_store_rows : 0); }

int cbk_build_doc_end_bits(const int64_t* d_pfxsum, int64_t n_docs, int64_t n_store_rows, uint32_t* d_bits, void* stream) {
  CBK_CHECK_ARG(d_pfxsum && d_bits, "cbk_build_doc_end_bits: null pointer argument");
  CBK_CHECK_ARG(n_docs > 0 && n_store_rows > 0, "cbk_build_doc_end_bits: sizes must be positive");
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return doc_end_bits_dispatch(d_pfxsum, n_docs, n_store_rows, d_bits, static_cast<cudaStream_t>(stream));
}

size_t cbk_maxsim_exhaustive_workspace_bytes(int64_t n_queries) { return exhaustive_workspace_bytes(n_queries > 0 ? n_queries : 1); }

int cbk_maxsim_exhaustive(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                          const uint32_t* d_doc_end_bits, int64_t n_docs, const int32_t* strides, int n_strides,
                          const float* d_Q, int q_len, int64_t n_queries, float* d_out_scores, void* d_workspace,
                          size_t workspace_bytes, int flags, void* stream) {
  CBK_CHECK_ARG(d_store && d_pfxsum && d_doc_end_bits && d_Q && d_out_scores, "cbk_maxsim_exhaustive: null pointer argument");
  CBK_CHECK_ARG(store_dtype == CBK_F16 || store_dtype == CBK_BF16, "cbk_maxsim_exhaustive: unknown store dtype %d", store_dtype);
  CBK_CHECK_ARG(n_store_rows > 0 && n_docs > 0 && n_queries > 0, "cbk_maxsim_exhaustive: sizes must be positive");
  CBK_CHECK_ARG(n_strides >= 0 && n_strides <= CBK_MAX_STRIDES && (n_strides == 0 || strides),
                "cbk_maxsim_exhaustive: n_strides %d outside [0, %d] or strides is NULL", n_strides, CBK_MAX_STRIDES);
  CBK_CHECK_SUPPORTED(dim == 128, "cbk_maxsim_exhaustive: dim %d not supported (128 only)", dim);
  CBK_CHECK_SUPPORTED(q_len >= 1 && q_len <= CBK_MAX_QLEN, "cbk_maxsim_exhaustive: q_len %d outside [1, %d]", q_len, CBK_MAX_QLEN);
  CBK_CHECK_SUPPORTED(n_store_rows < (1ll << 31) && n_queries < (1 << 20), "cbk_maxsim_exhaustive: store or batch too large");
  CBK_CHECK_ARG((reinterpret_cast<uintptr_t>(d_store) & 0xff) == 0, "cbk_maxsim_exhaustive: store base must be 256-byte aligned");
  if (!d_workspace || workspace_bytes < cbk_maxsim_exhaustive_workspace_bytes(n_queries) ||
      (reinterpret_cast<uintptr_t>(d_workspace) & 0xff) != 0) {
    set_error("cbk_maxsim_exhaustive: workspace of %zu bytes (256-byte aligned), need %zu", workspace_bytes,
              cbk_maxsim_exhaustive_workspace_bytes(n_queries));
    return CBK_ERR_WORKSPACE;
  }
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return exhaustive_dispatch(d_store, store_dtype, n_store_rows, d_pfxsum, d_doc_end_bits, n_docs, strides, n_strides, d_Q,
                             q_len, n_queries, d_out_scores, d_workspace, flags, static_cast<cudaStream_t>(stream));
}

size_t cbk_topk_dense_workspace_bytes(int64_t n_queries, int64_t n_docs, int k) {
  if (n_queries <= 0 || n_docs <= 0 || k <= 0) return 256;
  return topk_dense_workspace_bytes(n_queries, n_docs, k);
}

int cbk_topk_dense(const float* d_scores, int64_t n_queries, int64_t n_docs, int k, int64_t pid_base, int as_keys,
                   float* d_out_scores, int64_t* d_out_pids, void* d_workspace, size_t workspace_bytes, void* stream) {
  CBK_CHECK_ARG(d_scores && d_out_pids && (as_keys || d_out_scores), "cbk_topk_dense: null pointer argument");
  CBK_CHECK_ARG(n_queries > 0 && n_docs > 0 && k > 0, "cbk_topk_dense: sizes must be positive");
  CBK_CHECK_SUPPORTED(k <= topk_max_candidates() / 2, "cbk_topk_dense: k %d exceeds %lld", k, (long long)topk_max_candidates() / 2);
  CBK_CHECK_SUPPORTED(n_queries < 65536 && pid_base + n_docs < (1ll << 32), "cbk_topk_dense: batch or pid range too large");
  if (!d_workspace || workspace_bytes < cbk_topk_dense_workspace_bytes(n_queries, n_docs, k)) {
    set_error("cbk_topk_dense: workspace of %zu bytes, need %zu", workspace_bytes, cbk_topk_dense_workspace_bytes(n_queries, n_docs, k));
    return CBK_ERR_WORKSPACE;
  }
  int rc = check_device();
  if (rc != CBK_OK) return rc;
  return topk_dense_dispatch(d_scores, n_queries, n_docs, k, pid_base, as_keys, d_out_scores, d_out_pids, d_workspace,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
