// Candidate routing for a pid-range-sharded store (the reference's ranker is
// single-GPU, so this has no counterpart there): from the replicated CSR candidate lists keep, per
// query and in order, the candidates whose global pid lies in this shard's range [pid_lo, pid_hi).
// ONE launch (single pass with decoupled look-back): a CTA takes tiles of 32 queries in ticket order, counts the kept
// candidates of each (one warp per query, 4 queries per warp), publishes the tile's total, finds its global base by
// looking back over the totals / inclusive prefixes of the tiles before it, and scatters — the candidate ids are read
// from HBM once (the second read of a tile's 256 KB comes out of L1/L2).  Pure integer, HBM-bound work: 8 B read per
// candidate, 8 B written per kept candidate.  (The three-launch form — count, scan, scatter — is still used by the
// candidate-generation post-processing below.)
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

// ---- single-pass routing ---------------------------------------------------------------------------------
constexpr int kPartWarps = 8;
constexpr int kPartQPerWarp = 4;
constexpr int kPartTileQ = kPartWarps * kPartQPerWarp;      // queries per tile
constexpr uint64_t kPartAgg = 1ull << 62;                   // state = flag | value: tile total published
constexpr uint64_t kPartIncl = 2ull << 62;                  // inclusive prefix published
constexpr uint64_t kPartVal = (1ull << 62) - 1;

// state[0] = ticket counter, state[1 + t] = tile t's status word; zeroed by the caller before every launch
__global__ void __launch_bounds__(kPartWarps * 32)
partition_single_pass_kernel(const int64_t* __restrict__ pids, const int64_t* __restrict__ rowptr, int64_t n_queries,
                             int64_t lo, int64_t hi, int64_t* __restrict__ out_pids, int64_t* __restrict__ out_rowptr,
                             unsigned long long* __restrict__ state) {
  __shared__ int64_t s_cnt[kPartTileQ];
  __shared__ int64_t s_base;
  __shared__ int64_t s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (n_queries + kPartTileQ - 1) / kPartTileQ;
  while (true) {
    if (threadIdx.x == 0) s_tile = static_cast<int64_t>(atomicAdd(&state[0], 1ull));
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile >= n_tiles) break;
    const int64_t q0 = tile * kPartTileQ;
    // ---- count -----------------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < kPartQPerWarp; ++j) {
      const int64_t q = q0 + warp * kPartQPerWarp + j;
      int c = 0;
      if (q < n_queries) {
        const int64_t beg = rowptr[q], end = rowptr[q + 1];
        for (int64_t i = beg + lane; i < end; i += 32) {
          const int64_t p = pids[i];
          c += (p >= lo && p < hi) ? 1 : 0;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) s_cnt[warp * kPartQPerWarp + j] = c;
    }
    __syncthreads();
    // ---- tile scan + decoupled look-back (warp 0) -----------------------------------------------------------
    if (warp == 0) {
      const int64_t mine = s_cnt[lane];
      int64_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int64_t total = __shfl_sync(0xffffffffu, incl, 31);
      s_cnt[lane] = incl - mine;                              // exclusive offset of query `lane` inside the tile
      int64_t base = 0;
      if (lane == 0) {
        volatile unsigned long long* st = state + 1;
        if (tile > 0) {
          st[tile] = kPartAgg | static_cast<unsigned long long>(total);
          __threadfence();
          for (int64_t t = tile - 1; t >= 0; --t) {             // tiles are ticketed in order: predecessors are running
            unsigned long long w;
            do { w = st[t]; } while (w == 0ull);
            base += static_cast<int64_t>(w & kPartVal);
            if (w & kPartIncl) break;
          }
        }
        __threadfence();
        st[tile] = kPartIncl | static_cast<unsigned long long>(base + total);
        s_base = base;
        if (tile == n_tiles - 1) out_rowptr[n_queries] = base + total;
      }
    }
    __syncthreads();
    // ---- ordered scatter ----------------------------------------------------------------------------------
    const int64_t base = s_base;
#pragma unroll
    for (int j = 0; j < kPartQPerWarp; ++j) {
      const int64_t q = q0 + warp * kPartQPerWarp + j;
      if (q >= n_queries) continue;
      const int64_t beg = rowptr[q], end = rowptr[q + 1];
      int64_t dst = base + s_cnt[warp * kPartQPerWarp + j];
      if (lane == 0) out_rowptr[q] = dst;
      for (int64_t i0 = beg; i0 < end; i0 += 32) {
        const int64_t i = i0 + lane;
        const int64_t p = i < end ? pids[i] : -1;
        const bool keep = i < end && p >= lo && p < hi;
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        if (keep) out_pids[dst + __popc(m & ((1u << lane) - 1u))] = p;
        dst += __popc(m);
      }
    }
    __syncthreads();                                         // s_cnt / s_base / s_tile are reused by the next tile
  }
}

// ---- candidate-generation post-processing (reference colbert_ranker.py:163-174, 212-235) ----------------
// emb2pid[t] = document that owns store row t (ColbertIndex.build_emb2pid): one warp per document.
__global__ void __launch_bounds__(256)
build_emb2pid_kernel(const int64_t* __restrict__ pfxsum, int64_t n_docs, int32_t* __restrict__ emb2pid) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t d = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); d < n_docs; d += warps) {
    const int64_t beg = pfxsum[d], end = pfxsum[d + 1];
    for (int64_t t = beg + lane; t < end; t += 32) emb2pid[t] = static_cast<int32_t>(d);
  }
}

// One CTA per query: embedding ids → pids (emb2pid lookup) → sorted unique pids (the reference's per-query
// `list(set(...))`, ColbertIndex.embedding_ids_to_pids + uniq).  Ids outside [0, n_tokens) are dropped (faiss
// pads missing neighbours with -1).  Output: padded row of `n_ids` slots + count.
__global__ void __launch_bounds__(256)
unique_pids_kernel(const int64_t* __restrict__ emb_ids, int n_ids, const int32_t* __restrict__ emb2pid, int64_t n_tokens,
                   int P, int64_t* __restrict__ out_padded, int64_t* __restrict__ counts) {
  extern __shared__ uint32_t skeys[];          // P keys, then 8 warp totals
  __shared__ int warp_tot[8];
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < P; i += 256) {
    uint32_t key = 0xffffffffu;                // padding sorts last (ascending)
    if (i < n_ids) {
      const int64_t e = emb_ids[q * n_ids + i];
      if (e >= 0 && e < n_tokens) key = static_cast<uint32_t>(emb2pid[e]);
    }
    skeys[i] = key;
  }
  __syncthreads();
  // bitonic sort, ascending
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (P >> 1); t += 256) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const uint32_t a = skeys[lo], b = skeys[hi];
        if ((a > b) == asc) {
          skeys[lo] = b;
          skeys[hi] = a;
        }
      }
      __syncthreads();
    }
  // ordered compaction of first occurrences: each thread owns a contiguous run of P/256 keys
  const int per = P / 256 > 0 ? P / 256 : 1;
  const int beg = tid * per, end = min(P, beg + per);
  int mine = 0;
  for (int i = beg; i < end && tid * per < P; ++i) {
    const uint32_t k = skeys[i];
    if (k != 0xffffffffu && (i == 0 || skeys[i - 1] != k)) ++mine;
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  int pos = base + incl - mine;
  for (int i = beg; i < end && tid * per < P; ++i) {
    const uint32_t k = skeys[i];
    if (k != 0xffffffffu && (i == 0 || skeys[i - 1] != k)) out_padded[q * n_ids + pos++] = static_cast<int64_t>(k);
  }
  if (tid == 255) counts[q] = base + incl;
}

// padded rows + rowptr → CSR
__global__ void __launch_bounds__(256)
pack_rows_kernel(const int64_t* __restrict__ padded, int n_ids, const int64_t* __restrict__ rowptr, int64_t n_queries,
                 int64_t* __restrict__ out) {
  const int64_t q = blockIdx.x;
  const int64_t beg = rowptr[q], n = rowptr[q + 1] - beg;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[beg + i] = padded[q * n_ids + i];
}

}  // namespace

int emb2pid_dispatch(const int64_t* d_pfxsum, int64_t n_docs, int32_t* d_emb2pid, cudaStream_t stream) {
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n_docs + 7) / 8, static_cast<int64_t>(sm_count()) * 16)));
  build_emb2pid_kernel<<<grid, 256, 0, stream>>>(d_pfxsum, n_docs, d_emb2pid);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

int unique_pids_dispatch(const int64_t* d_emb_ids, int64_t n_queries, int n_ids, const int32_t* d_emb2pid, int64_t n_tokens,
                         int64_t* d_out_pids, int64_t* d_out_rowptr, void* d_workspace, cudaStream_t stream) {
  int P = 256;
  while (P < n_ids) P <<= 1;
  int64_t* padded = static_cast<int64_t*>(d_workspace);
  int64_t* counts = padded + n_queries * n_ids;
  const size_t smem = static_cast<size_t>(P) * sizeof(uint32_t);
  CBK_CUDA(cudaFuncSetAttribute(unique_pids_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  unique_pids_kernel<<<static_cast<unsigned int>(n_queries), 256, smem, stream>>>(d_emb_ids, n_ids, d_emb2pid, n_tokens, P,
                                                                                  padded, counts);
  CBK_CUDA(cudaGetLastError());
  exclusive_scan_kernel<<<1, 1024, 0, stream>>>(counts, n_queries, d_out_rowptr);
  CBK_CUDA(cudaGetLastError());
  pack_rows_kernel<<<static_cast<unsigned int>(n_queries), 256, 0, stream>>>(padded, n_ids, d_out_rowptr, n_queries, d_out_pids);
  CBK_CUDA(cudaGetLastError());
  count_launch(3);
  return CBK_OK;
}

size_t partition_workspace_bytes(int64_t n_queries) {
  const int64_t n_tiles = (std::max<int64_t>(n_queries, 1) + kPartTileQ - 1) / kPartTileQ;
  return static_cast<size_t>(n_tiles + 1) * sizeof(unsigned long long);
}

int partition_dispatch(const int64_t* d_pids, const int64_t* d_rowptr, int64_t n_queries, int64_t lo, int64_t hi,
                       int64_t* d_out_pids, int64_t* d_out_rowptr, void* d_workspace, cudaStream_t stream) {
  const int64_t n_tiles = (n_queries + kPartTileQ - 1) / kPartTileQ;
  CBK_CUDA(cudaMemsetAsync(d_workspace, 0, partition_workspace_bytes(n_queries), stream));
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_tiles, static_cast<int64_t>(sm_count()) * 8)));
  partition_single_pass_kernel<<<grid, kPartWarps * 32, 0, stream>>>(d_pids, d_rowptr, n_queries, lo, hi, d_out_pids, d_out_rowptr,
                                                                    static_cast<unsigned long long*>(d_workspace));
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk
