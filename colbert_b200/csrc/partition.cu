// Candidate routing for a pid-range-sharded store (the reference's ranker is
// single-GPU, so this has no counterpart there): from the replicated CSR candidate lists keep, per
// query and in order, the candidates whose global pid lies in this shard's range [pid_lo, pid_hi).
// Three tiny launches: count per query (one warp per query) → exclusive scan of the counts →
// ordered scatter (one warp per query, ballot compaction).  Pure integer, HBM-bound work:
// 8 B read per candidate twice, 8 B written per kept candidate.
#include <algorithm>

#include "cbk_common.cuh"

namespace cbk {

namespace {

__global__ void __launch_bounds__(256)
count_in_range_kernel(const int64_t* __restrict__ pids, const int64_t* __restrict__ rowptr, int64_t n_queries,
                      int64_t lo, int64_t hi, int64_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); q < n_queries; q += warps) {
    const int64_t beg = rowptr[q], end = rowptr[q + 1];
    int c = 0;
    for (int64_t i = beg + lane; i < end; i += 32) {
      const int64_t p = pids[i];
      c += (p >= lo && p < hi) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[q] = c;
  }
}

// out[0] = 0, out[i+1] = counts[0] + … + counts[i]; single CTA, chunked Hillis-Steele over 1024 lanes.
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(const int64_t* __restrict__ counts, int64_t n, int64_t* __restrict__ out) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    carry = 0;
    out[0] = 0;
  }
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + tid;
    int64_t v = i < n ? counts[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int64_t prefix = carry + (warp > 0 ? warp_sums[warp - 1] : 0);
    if (i < n) out[i + 1] = prefix + v;
    __syncthreads();
    if (tid == 1023) carry = prefix + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
scatter_in_range_kernel(const int64_t* __restrict__ pids, const int64_t* __restrict__ rowptr, int64_t n_queries,
                        int64_t lo, int64_t hi, const int64_t* __restrict__ out_rowptr, int64_t* __restrict__ out_pids) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); q < n_queries; q += warps) {
    const int64_t beg = rowptr[q], end = rowptr[q + 1];
    int64_t dst = out_rowptr[q];
    for (int64_t i0 = beg; i0 < end; i0 += 32) {
      const int64_t i = i0 + lane;
      const int64_t p = i < end ? pids[i] : -1;
      const bool keep = i < end && p >= lo && p < hi;
      const unsigned int m = __ballot_sync(0xffffffffu, keep);
      if (keep) out_pids[dst + __popc(m & ((1u << lane) - 1u))] = p;
      dst += __popc(m);
    }
  }
}

}  // namespace

int partition_dispatch(const int64_t* d_pids, const int64_t* d_rowptr, int64_t n_queries, int64_t lo, int64_t hi,
                       int64_t* d_out_pids, int64_t* d_out_rowptr, int64_t* d_counts, cudaStream_t stream) {
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n_queries + 7) / 8, static_cast<int64_t>(sm_count()) * 8)));
  count_in_range_kernel<<<grid, 256, 0, stream>>>(d_pids, d_rowptr, n_queries, lo, hi, d_counts);
  CBK_CUDA(cudaGetLastError());
  exclusive_scan_kernel<<<1, 1024, 0, stream>>>(d_counts, n_queries, d_out_rowptr);
  CBK_CUDA(cudaGetLastError());
  scatter_in_range_kernel<<<grid, 256, 0, stream>>>(d_pids, d_rowptr, n_queries, lo, hi, d_out_rowptr, d_out_pids);
  CBK_CUDA(cudaGetLastError());
  count_launch(3);
  return CBK_OK;
}

}  // namespace cbk
