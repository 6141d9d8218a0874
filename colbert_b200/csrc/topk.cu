// Per-query top-k (sm_100a).  Replaces reference colbert/ranking/colbert_ranker.py:128-130
// (full descending sort of the candidate scores, truncated to `depth`).
//
// One CTA per query.  Each candidate becomes one 64-bit key
//     [ ordered(score) : 32 | ~pid : 32 ]
// so that a single descending sort realises the total order (score desc, pid asc).  Keys are
// sorted by a bitonic network held in shared memory; the compare-exchange distances below 32 run
// inside a warp on registers with shuffles (no block barrier), the larger ones through smem.
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

// keys sorted DESCENDING.  P = padded power-of-two length (≥ 32).
__global__ void __launch_bounds__(kTopkThreads)
topk_per_query_kernel(const float* __restrict__ scores, const int64_t* __restrict__ cand_pids,
                      const int64_t* __restrict__ rowptr, int P, int k, float* __restrict__ out_scores,
                      int64_t* __restrict__ out_pids) {
  extern __shared__ uint64_t keys[];
  const int64_t q = blockIdx.x;
  const int64_t beg = rowptr[q];
  const int n = static_cast<int>(min(rowptr[q + 1] - beg, static_cast<int64_t>(P)));
  const int tid = threadIdx.x;

  for (int i = tid; i < P; i += kTopkThreads) {
    uint64_t key = 0;  // padding: below every real key
    if (i < n) {
      const uint32_t s = float_to_ordered(scores[beg + i] + 0.0f);  // -0.0 → +0.0: they tie
      const uint32_t p = ~static_cast<uint32_t>(cand_pids[beg + i]);
      key = (static_cast<uint64_t>(s) << 32) | p;
    }
    keys[i] = key;
  }
  __syncthreads();

  for (int size = 2; size <= P; size <<= 1) {
    int stride = size >> 1;
    // distances ≥ 32: one compare-exchange per pair through shared memory
    for (; stride >= 32; stride >>= 1) {
      for (int t = tid; t < (P >> 1); t += kTopkThreads) {
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
    for (int base = (tid >> 5) * 32; base < P; base += kTopkThreads) {
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

  for (int i = tid; i < k; i += kTopkThreads) {
    float s = -INFINITY;
    int64_t pid = -1;
    if (i < n) {
      const uint64_t key = keys[i];
      s = ordered_to_float(static_cast<uint32_t>(key >> 32));
      pid = static_cast<int64_t>(~static_cast<uint32_t>(key));
    }
    out_scores[q * k + i] = s;
    out_pids[q * k + i] = pid;
  }
}

}  // namespace

int64_t topk_max_candidates() { return kTopkMaxCand; }

int topk_dispatch(const float* d_scores, const int64_t* d_cand_pids, const int64_t* d_cand_rowptr, int64_t n_queries,
                  int64_t max_cand_per_query, int k, float* d_out_scores, int64_t* d_out_pids, cudaStream_t stream) {
  int P = 32;
  while (P < max_cand_per_query) P <<= 1;
  const size_t smem = static_cast<size_t>(P) * sizeof(uint64_t);
  CBK_CUDA(cudaFuncSetAttribute(topk_per_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  topk_per_query_kernel<<<static_cast<unsigned int>(n_queries), kTopkThreads, smem, stream>>>(
      d_scores, d_cand_pids, d_cand_rowptr, P, k, d_out_scores, d_out_pids);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk
