// Per-query top-k (sm_100a).  Replaces reference colbert/ranking/colbert_ranker.py:128-130
// (full descending sort of the candidate scores, truncated to `depth`).
//
// One CTA per query.  Each candidate becomes one 64-bit key
//     [ ordered(score) : 32 | ~pid : 32 ]
// so that a single descending sort realises the total order (score desc, pid asc).  Keys are
// sorted by a bitonic network held in shared memory; the compare-exchange distances below 32 run
// inside a warp on registers with shuffles (no block barrier), the larger ones through smem.
#include <algorithm>

#include "cbk_common.cuh"

namespace cbk {

namespace {

constexpr int kTopkThreads = 256;
constexpr int kTopkMaxCand = 16384;  // the reference's BSIZE (colbert_ranker.py:11)

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
  uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, mask);
  hi = __shfl_xor_sync(0xffffffffu, hi, mask);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Sort the P keys held in shared memory, descending.  P = power of two ≥ 32.
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int P, int tid) {
  const int nthreads = blockDim.x;
  for (int size = 2; size <= P; size <<= 1) {
    int stride = size >> 1;
    // distances ≥ 32: one compare-exchange per pair through shared memory
    for (; stride >= 32; stride >>= 1) {
      for (int t = tid; t < (P >> 1); t += nthreads) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
    // distances < 32: each warp owns 32 consecutive keys per pass, in registers
    for (int base = (tid >> 5) * 32; base < P; base += nthreads) {
      const int i = base + (tid & 31);
      uint64_t v = keys[i];
      const bool desc = (i & size) == 0;
      for (int s = min(stride, 16); s >= 1; s >>= 1) {
        const uint64_t o = shfl_xor_u64(v, s);
        const bool lower = (i & s) == 0;
        // in a descending run the lower index keeps the larger key
        const bool keep_max = (lower == desc);
        v = keep_max ? (v > o ? v : o) : (v < o ? v : o);
      }
      keys[i] = v;
    }
    __syncthreads();
  }
}

// key 0 is padding (sorts below every real key): decoded as (−inf, −1)
__device__ __forceinline__ void emit(uint64_t key, int64_t pos, float* out_scores, int64_t* out_pids,
                                     uint64_t* out_keys) {
  if (out_keys) {
    out_keys[pos] = key;
  } else {
    out_scores[pos] = key ? ordered_to_float(static_cast<uint32_t>(key >> 32)) : -INFINITY;
    out_pids[pos] = key ? static_cast<int64_t>(~static_cast<uint32_t>(key)) : -1;
  }
}

// One CTA per query: candidates (scores + pids, CSR) → top-k as (scores, pids) or as packed keys.
__global__ void __launch_bounds__(kTopkThreads)
topk_per_query_kernel(const float* __restrict__ scores, const int64_t* __restrict__ cand_pids,
                      const int64_t* __restrict__ rowptr, int Pmax, int k, int neg_inf_is_padding,
                      float* __restrict__ out_scores, int64_t* __restrict__ out_pids,
                      uint64_t* __restrict__ out_keys) {
  extern __shared__ uint64_t keys[];
  const int64_t q = blockIdx.x;
  const int64_t beg = rowptr[q];
  const int n = static_cast<int>(min(rowptr[q + 1] - beg, static_cast<int64_t>(Pmax)));
  const int tid = threadIdx.x;
  int P = 32;  // sort only as many slots as this query needs (routed lists are short)
  while (P < n) P <<= 1;
  for (int i = tid; i < P; i += kTopkThreads) {
    uint64_t key = 0;
    if (i < n) {
      const float sc = scores[beg + i] + 0.0f;  // -0.0 → +0.0: they tie
      if (!(neg_inf_is_padding && sc == -INFINITY)) {
        const uint32_t p = ~static_cast<uint32_t>(cand_pids[beg + i]);
        key = (static_cast<uint64_t>(float_to_ordered(sc)) << 32) | p;
      }
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc(keys, P, tid);
  for (int i = tid; i < k; i += kTopkThreads) emit(i < n ? keys[i] : 0ull, q * k + i, out_scores, out_pids, out_keys);
}

// Merge of sorted key lists.  CTA (g, q) merges lists g*G .. g*G+G-1 of query q; list w of query q starts at
// in_keys[w*sw + q*sq].  Covers both layouts in use: rank-major [W, n_queries, k_in] as all-gathered from the
// shards (sw = n_queries*k_in, sq = k_in) and query-major [n_queries, n_lists, k_in] between the
// levels of the exhaustive top-k (sw = k_in, sq = n_lists*k_in).  The key order is a total order, so the
// result does not depend on how the lists were split.  Output: keys (intermediate level) or decoded
// (scores, pids) (last level), at row q*gridDim.x + g.
__global__ void __launch_bounds__(1024)
merge_topk_keys_kernel(const uint64_t* __restrict__ in_keys, int n_lists, int G, int64_t sw, int64_t sq, int k_in, int Pmax,
                       int k, float* __restrict__ out_scores, int64_t* __restrict__ out_pids,
                       uint64_t* __restrict__ out_keys) {
  extern __shared__ uint64_t keys[];
  const int g = blockIdx.x;
  const int64_t q = blockIdx.y;
  const int first = g * G;
  const int cnt = min(G, n_lists - first);
  const int n = cnt * k_in;
  const int tid = threadIdx.x;
  int P = 32;
  while (P < n) P <<= 1;
  for (int i = tid; i < P; i += blockDim.x) {
    uint64_t key = 0;
    if (i < n) {
      const int w = i / k_in, j = i - w * k_in;
      key = in_keys[static_cast<int64_t>(first + w) * sw + q * sq + j];
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc(keys, P, tid);
  const int64_t row = q * gridDim.x + g;
  for (int i = tid; i < k; i += blockDim.x) emit(i < n ? keys[i] : 0ull, row * k + i, out_scores, out_pids, out_keys);
}

// Dense score rows (exhaustive scoring): CTA (c, q) sorts chunk c of row q of scores[n_queries, n_docs]
// (pid = pid_base + column) and emits its top-k keys at out_keys[(q*gridDim.x + c)*k ..].
__global__ void __launch_bounds__(1024)
topk_dense_chunks_kernel(const float* __restrict__ scores, int64_t n_docs, int chunk, int64_t pid_base, int k,
                         uint64_t* __restrict__ out_keys) {
  extern __shared__ uint64_t keys[];
  const int c = blockIdx.x;
  const int64_t q = blockIdx.y;
  const int64_t first = static_cast<int64_t>(c) * chunk;
  const int n = static_cast<int>(min(static_cast<int64_t>(chunk), n_docs - first));
  const int tid = threadIdx.x;
  int P = 32;
  while (P < n) P <<= 1;
  const float* row = scores + q * n_docs + first;
  for (int i = tid; i < P; i += blockDim.x) {
    uint64_t key = 0;
    if (i < n) {
      const uint32_t p = ~static_cast<uint32_t>(pid_base + first + i);
      key = (static_cast<uint64_t>(float_to_ordered(row[i] + 0.0f)) << 32) | p;
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort_desc(keys, P, tid);
  const int64_t orow = q * gridDim.x + c;
  for (int i = tid; i < k; i += blockDim.x) out_keys[orow * k + i] = i < n ? keys[i] : 0ull;
}

}  // namespace

int64_t topk_max_candidates() { return kTopkMaxCand; }

static int padded_pow2(int64_t n) {
  int P = 32;
  while (P < n) P <<= 1;
  return P;
}

int topk_dispatch(const float* d_scores, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr, int64_t n_queries,
                  int64_t max_cand_per_query, int k, int flags, float* d_out_scores, int64_t* d_out_pids,
                  uint64_t* d_out_keys, cudaStream_t stream) {
  const int P = padded_pow2(max_cand_per_query);
  const size_t smem = static_cast<size_t>(P) * sizeof(uint64_t);
  CBK_CUDA(cudaFuncSetAttribute(topk_per_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  topk_per_query_kernel<<<static_cast<unsigned int>(n_queries), kTopkThreads, smem, stream>>>(
      d_scores, d_cand_pids, d_cand_rowptr, P, k, (flags & CBK_TOPK_NEG_INF_IS_PADDING) ? 1 : 0, d_out_scores,
      d_out_pids, d_out_keys);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

static int launch_merge(const uint64_t* in, int n_lists, int G, int64_t sw, int64_t sq, int k_in, int64_t n_queries, int k,
                        float* out_scores, int64_t* out_pids, uint64_t* out_keys, cudaStream_t stream) {
  const int n_groups = (n_lists + G - 1) / G;
  const int P = padded_pow2(static_cast<int64_t>(std::min(G, n_lists)) * k_in);
  const size_t smem = static_cast<size_t>(P) * sizeof(uint64_t);
  const int threads = P >= 4096 ? 1024 : kTopkThreads;
  CBK_CUDA(cudaFuncSetAttribute(merge_topk_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  merge_topk_keys_kernel<<<dim3(n_groups, static_cast<unsigned int>(n_queries)), threads, smem, stream>>>(
      in, n_lists, G, sw, sq, k_in, P, k, out_scores, out_pids, out_keys);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

int merge_dispatch(const uint64_t* d_keys, int world, int64_t n_queries, int k_in, int k, float* d_out_scores,
                   int64_t* d_out_pids, cudaStream_t stream) {
  return launch_merge(d_keys, world, world, n_queries * k_in, k_in, k_in, n_queries, k, d_out_scores, d_out_pids, nullptr,
                      stream);
}

// ---- top-k over dense score rows [n_queries, n_docs] (exhaustive scoring) ----------------------------
constexpr int kDenseChunk = 16384;

size_t topk_dense_workspace_bytes(int64_t n_queries, int64_t n_docs, int k) {
  const int64_t nch = (n_docs + kDenseChunk - 1) / kDenseChunk;
  const int64_t k1 = std::min<int64_t>(k, kDenseChunk);
  const int64_t G = std::max<int64_t>(2, kTopkMaxCand / k1);
  const int64_t lvl2 = (nch + G - 1) / G;
  return static_cast<size_t>(n_queries * (nch + lvl2) * k1) * sizeof(uint64_t) + 256;
}

int topk_dense_dispatch(const float* d_scores, int64_t n_queries, int64_t n_docs, int k, int64_t pid_base, int as_keys,
                        float* d_out_scores, int64_t* d_out_pids, void* d_workspace, cudaStream_t stream) {
  const int nch = static_cast<int>((n_docs + kDenseChunk - 1) / kDenseChunk);
  const int k1 = static_cast<int>(std::min<int64_t>(k, std::min<int64_t>(kDenseChunk, n_docs)));
  const int G = std::max(2, kTopkMaxCand / k1);
  uint64_t* bufA = static_cast<uint64_t*>(d_workspace);
  uint64_t* bufB = bufA + n_queries * nch * k1;
  const int P = padded_pow2(std::min<int64_t>(kDenseChunk, n_docs));
  const size_t smem = static_cast<size_t>(P) * sizeof(uint64_t);
  CBK_CUDA(cudaFuncSetAttribute(topk_dense_chunks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  topk_dense_chunks_kernel<<<dim3(nch, static_cast<unsigned int>(n_queries)), P >= 4096 ? 1024 : kTopkThreads, smem, stream>>>(
      d_scores, n_docs, kDenseChunk, pid_base, k1, bufA);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  int n_lists = nch;
  uint64_t* in = bufA;
  uint64_t* out = bufB;
  while (true) {
    const int n_groups = (n_lists + G - 1) / G;
    const bool last = n_groups == 1;
    int rc = launch_merge(in, n_lists, G, k1, static_cast<int64_t>(n_lists) * k1, k1, n_queries, last ? k : k1,
                          (last && !as_keys) ? d_out_scores : nullptr, (last && !as_keys) ? d_out_pids : nullptr,
                          last ? (as_keys ? reinterpret_cast<uint64_t*>(d_out_pids) : nullptr) : out, stream);
    if (rc != CBK_OK) return rc;
    if (last) break;
    n_lists = n_groups;
    std::swap(in, out);
  }
  return CBK_OK;
}

}  // namespace cbk
