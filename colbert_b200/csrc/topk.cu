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

// ---- radix select over dense score rows --------------------------------------------------------------
// Top-k of a long row (exhaustive scoring: up to millions of documents per query) without sorting it: three
// histogram levels over the 32-bit ordered score (12 + 12 + 8 bits) locate the k-th largest score; everything above
// it plus the bin that holds it is compacted (≤ kTopkMaxCand keys, normally ≈ k) and only that is sorted.  A level is
// skipped as soon as the survivors fit.  The score matrix is read three times instead of being bitonic-sorted in
// 16384-wide chunks (2.0 → 0.2 ms for 16 × 1.1 M scores, k = 1000).
constexpr int kSelChunk = 16384;          // scores per CTA at most; fewer when the batch is small (sel_chunk): a single query over
                                          // 1.1 M documents would otherwise occupy 68 of 148 SMs
constexpr int kSelThreads = 256;
constexpr int kSelBins = 4096;

struct SelState {            // per query, zero-initialised per call
  uint32_t prefix;           // resolved high bits of the k-th largest ordered score
  uint32_t bits;             // how many bits are resolved (0, 12, 24, 32)
  uint32_t k_rem;            // how many of the top k lie inside the current bin
  uint32_t done;             // survivors (above + inside the bin) fit into kTopkMaxCand: no further level
  uint32_t n_cand;           // compaction cursor
  uint32_t overflow;         // more than kTopkMaxCand scores tie with the k-th: this query takes the chunk-sort path
  uint32_t pad[2];
};

__device__ __forceinline__ uint32_t sel_key(float x) { return float_to_ordered(x + 0.0f); }   // -0.0 ≡ +0.0

// level L histogram of the scores whose resolved prefix matches; hist [n_queries][kSelBins]
template <int kLevel>
__global__ void __launch_bounds__(kSelThreads)
radix_hist_kernel(const float* __restrict__ scores, int64_t n_docs, int chunk, const SelState* __restrict__ state,
                  uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kSelBins];
  const int64_t q = blockIdx.y;
  const SelState stq = state[q];
  if (kLevel > 0 && stq.done) return;
  constexpr int kShift = kLevel == 0 ? 20 : (kLevel == 1 ? 8 : 0);
  constexpr uint32_t kMask = kLevel == 2 ? 0xffu : 0xfffu;
  constexpr int kPrevShift = kLevel == 1 ? 20 : 8;       // (unused at level 0: nothing is resolved yet)
  for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) sh[i] = 0;
  __syncthreads();
  const int64_t first = static_cast<int64_t>(blockIdx.x) * chunk;
  const int n = static_cast<int>(min(static_cast<int64_t>(chunk), n_docs - first));
  const float* row = scores + q * n_docs + first;
  for (int i0 = 0; i0 < n; i0 += kSelThreads) {
    const int i = i0 + threadIdx.x;
    bool take = i < n;
    uint32_t key = 0;
    if (take) {
      key = sel_key(row[i]);
      if (kLevel > 0) take = (key >> kPrevShift) == stq.prefix;
    }
    const uint32_t bin = (key >> kShift) & kMask;
    // scores cluster in a few bins (one binade = 8 level-0 bins): aggregate equal bins inside the warp first
    const uint32_t act = __ballot_sync(0xffffffffu, take);
    if (take) {
      const uint32_t peers = __match_any_sync(act, bin);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[bin], __popc(peers));
    }
  }
  __syncthreads();
  uint32_t* hq = hist + q * kSelBins;
  for (int i = threadIdx.x; i < kSelBins; i += kSelThreads)
    if (sh[i]) atomicAdd(&hq[i], sh[i]);
}

// one CTA per query: walk the level histogram from the top until k_rem scores are covered
template <int kLevel>
__global__ void __launch_bounds__(kSelThreads)
radix_select_kernel(uint32_t* __restrict__ hist, SelState* __restrict__ state, int k) {
  __shared__ uint32_t part[kSelThreads];
  __shared__ uint32_t s_bin, s_above;
  const int64_t q = blockIdx.x;
  SelState stq = state[q];
  if (kLevel > 0 && stq.done) return;
  if (kLevel == 0) stq.k_rem = static_cast<uint32_t>(k);
  constexpr int kBins = kLevel == 2 ? 256 : kSelBins;
  constexpr int kPer = kBins / kSelThreads;              // bins per thread (16 or 1)
  uint32_t* hq = hist + q * kSelBins;
  // thread t owns bins [hi - kPer + 1, hi] counted from the top: t = 0 holds the largest bins
  const int top = kBins - 1 - threadIdx.x * kPer;
  uint32_t mine = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) mine += hq[top - j];
  part[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t cum = 0;
    int t = 0;
    for (; t < kSelThreads; ++t) {
      if (cum + part[t] >= stq.k_rem) break;
      cum += part[t];
    }
    // k_rem ≤ the number of scores that reached this level, so the walk ends inside the histogram; the clamps only
    // keep a violated precondition from reading out of bounds
    t = min(t, kSelThreads - 1);
    int b = kBins - 1 - t * kPer;
    for (int j = 0; j < kPer - 1; ++j, --b) {
      if (cum + hq[b] >= stq.k_rem) break;
      cum += hq[b];
    }
    s_bin = static_cast<uint32_t>(b);
    s_above = cum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    constexpr int kBitsHere = kLevel == 2 ? 8 : 12;
    const uint32_t in_bin = hq[s_bin];
    stq.prefix = (stq.prefix << kBitsHere) | s_bin;
    stq.bits += kBitsHere;
    stq.k_rem -= s_above;
    const uint32_t n_above_total = static_cast<uint32_t>(k) - stq.k_rem;      // strictly above the bin, all levels
    // stop refining once the survivors are cheap to sort: one more histogram pass over the row costs less than
    // sorting 16384 keys in one CTA (0.17 ms), so the early levels only stop at ~4 k survivors
    const uint32_t limit = kLevel == 2 ? static_cast<uint32_t>(kTopkMaxCand)
                                       : min(static_cast<uint32_t>(kTopkMaxCand), max(4096u, 2u * static_cast<uint32_t>(k)));
    if (n_above_total + in_bin <= limit) stq.done = 1;
    else if (kLevel == 2) stq.overflow = 1;                                    // > kTopkMaxCand exact ties with the k-th
    state[q] = stq;
  }
}

// every score at or above the located bin becomes a packed key in cand[q][0 .. n_cand)
__global__ void __launch_bounds__(kSelThreads)
radix_compact_kernel(const float* __restrict__ scores, int64_t n_docs, int chunk, int64_t pid_base, SelState* __restrict__ state,
                     uint64_t* __restrict__ cand) {
  const int64_t q = blockIdx.y;
  const SelState stq = state[q];
  if (stq.overflow) return;
  const int shift = 32 - static_cast<int>(stq.bits);
  const int64_t first = static_cast<int64_t>(blockIdx.x) * chunk;
  const int n = static_cast<int>(min(static_cast<int64_t>(chunk), n_docs - first));
  const float* row = scores + q * n_docs + first;
  uint64_t* cq = cand + q * kTopkMaxCand;
  for (int i0 = 0; i0 < n; i0 += kSelThreads) {
    const int i = i0 + threadIdx.x;
    uint32_t key = 0;
    bool take = false;
    if (i < n) {
      key = sel_key(row[i]);
      take = shift >= 32 ? true : (key >> shift) >= stq.prefix;
    }
    const uint32_t act = __ballot_sync(0xffffffffu, take);
    if (act) {
      uint32_t base = 0;
      const int lane = threadIdx.x & 31;
      if (lane == __ffs(act) - 1) base = atomicAdd(&state[q].n_cand, __popc(act));
      base = __shfl_sync(0xffffffffu, base, __ffs(act) - 1);
      if (take) {
        const uint32_t pos = base + __popc(act & ((1u << lane) - 1u));
        if (pos < static_cast<uint32_t>(kTopkMaxCand))
          cq[pos] = (static_cast<uint64_t>(key) << 32) | ~static_cast<uint32_t>(pid_base + first + i);
      }
    }
  }
}

// one CTA per query: sort the survivors, emit the top k
__global__ void __launch_bounds__(1024)
radix_final_kernel(const uint64_t* __restrict__ cand, const SelState* __restrict__ state, int k, float* __restrict__ out_scores,
                   int64_t* __restrict__ out_pids, uint64_t* __restrict__ out_keys) {
  extern __shared__ uint64_t keys[];
  const int64_t q = blockIdx.x;
  const SelState stq = state[q];
  if (stq.overflow) return;                       // the chunk-sort path writes this query's row
  const int n = static_cast<int>(min(stq.n_cand, static_cast<uint32_t>(kTopkMaxCand)));
  const int tid = threadIdx.x;
  int P = 32;
  while (P < n) P <<= 1;
  const uint64_t* cq = cand + q * kTopkMaxCand;
  for (int i = tid; i < P; i += blockDim.x) keys[i] = i < n ? cq[i] : 0ull;
  __syncthreads();
  bitonic_sort_desc(keys, P, tid);
  for (int i = tid; i < k; i += blockDim.x) emit(i < n ? keys[i] : 0ull, q * k + i, out_scores, out_pids, out_keys);
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

// Short lists, shallow depth (n ≤ 1024, k ≤ 32 — the reference's own call: 1000 candidates, depth 10): a tournament
// in registers instead of a full sort.  The list is taken 256 keys at a time: every warp sorts its 32 keys with shuffles;
// then, three times, the upper half of the surviving warps hand their list (reversed) to the lower half, which keeps
// max(a[i], b[31-i]) — the top 32 of both, as a bitonic sequence — and merges it in five more shuffle steps; warp 0 folds
// the chunk's top 32 into the running top 32 the same way.  256 threads per query: eight CTAs per SM stay resident, and a
// routed list of ~125 candidates (8-GPU shards) costs one chunk.
constexpr int kSmallThreads = 256;

__device__ __forceinline__ uint64_t warp_merge32_desc(uint64_t v, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint64_t o = shfl_xor_u64(v, s);
    v = ((lane & s) == 0) ? (v > o ? v : o) : (v < o ? v : o);
  }
  return v;
}

// the tournament itself: key_of(i) yields key i of this CTA's list (0 = padding), n ≤ 1024; warp 0 returns the top 32
template <typename KeyOf>
__device__ __forceinline__ uint64_t tournament_top32(KeyOf key_of, int n, uint64_t (*ex)[32]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t run = 0;                                   // warp 0: the top 32 so far, descending along the lanes
  for (int base = 0; base < n || base == 0; base += kSmallThreads) {
    const int i = base + tid;
    uint64_t v = i < n ? key_of(i) : 0ull;
    // 32 keys per warp, sorted descending along the lanes
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
      for (int s = size >> 1; s >= 1; s >>= 1) {
        const uint64_t o = shfl_xor_u64(v, s);
        const bool keep_max = ((lane & s) == 0) == ((lane & size) == 0);
        v = keep_max ? (v > o ? v : o) : (v < o ? v : o);
      }
    const int n_warps = (min(n - base, kSmallThreads) + 31) >> 5;   // warps beyond hold only padding
    int active = 1;
    while (active < n_warps) active <<= 1;
    for (; active > 1; active >>= 1) {
      const int half = active >> 1;
      if (warp >= half && warp < active) ex[warp - half][31 - lane] = v;
      __syncthreads();
      if (warp < half) {
        const uint64_t o = ex[warp][lane];
        v = warp_merge32_desc(v > o ? v : o, lane);
      }
      __syncthreads();
    }
    if (warp == 0) {
      const uint64_t rv = shfl_xor_u64(v, 31);        // the chunk's top 32, reversed
      run = warp_merge32_desc(run > rv ? run : rv, lane);
    }
  }
  return run;
}

__global__ void __launch_bounds__(kSmallThreads)
topk_small_kernel(const float* __restrict__ scores, const int64_t* __restrict__ cand_pids, const int64_t* __restrict__ rowptr,
                  int k, int neg_inf_is_padding, float* __restrict__ out_scores, int64_t* __restrict__ out_pids,
                  uint64_t* __restrict__ out_keys) {
  __shared__ uint64_t ex[kSmallThreads / 64][32];
  const int64_t q = blockIdx.x;
  const int64_t beg = rowptr[q];
  const int n = static_cast<int>(min(rowptr[q + 1] - beg, static_cast<int64_t>(1024)));
  const uint64_t run = tournament_top32(
      [&](int i) -> uint64_t {
        const float sc = scores[beg + i] + 0.0f;      // -0.0 → +0.0: they tie
        if (neg_inf_is_padding && sc == -INFINITY) return 0ull;
        return (static_cast<uint64_t>(float_to_ordered(sc)) << 32) | ~static_cast<uint32_t>(cand_pids[beg + i]);
      },
      n, ex);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32 && lane < k) emit(run, q * k + lane, out_scores, out_pids, out_keys);
}

// the same tournament over already-packed keys: W short sorted lists per query (the all-gathered shard winners)
__global__ void __launch_bounds__(kSmallThreads)
merge_small_kernel(const uint64_t* __restrict__ in_keys, int n_lists, int64_t sw, int64_t sq, int k_in, int k,
                   float* __restrict__ out_scores, int64_t* __restrict__ out_pids, uint64_t* __restrict__ out_keys) {
  __shared__ uint64_t ex[kSmallThreads / 64][32];
  const int64_t q = blockIdx.x;
  const uint64_t run = tournament_top32(
      [&](int i) -> uint64_t {
        const int w = i / k_in, j = i - w * k_in;
        return in_keys[static_cast<int64_t>(w) * sw + q * sq + j];
      },
      n_lists * k_in, ex);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32 && lane < k) emit(run, q * k + lane, out_scores, out_pids, out_keys);
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
                       uint64_t* __restrict__ out_keys, const SelState* __restrict__ gate) {
  extern __shared__ uint64_t keys[];
  const int g = blockIdx.x;
  const int64_t q = blockIdx.y;
  if (gate && !gate[q].overflow) return;   // fallback of the radix-select path: only for the queries that need it
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
                         uint64_t* __restrict__ out_keys, const SelState* __restrict__ gate) {
  extern __shared__ uint64_t keys[];
  const int c = blockIdx.x;
  const int64_t q = blockIdx.y;
  if (gate && !gate[q].overflow) return;
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
  if (max_cand_per_query <= 1024 && k <= 32) {
    topk_small_kernel<<<static_cast<unsigned int>(n_queries), kSmallThreads, 0, stream>>>(
        d_scores, d_cand_pids, d_cand_rowptr, k, (flags & CBK_TOPK_NEG_INF_IS_PADDING) ? 1 : 0, d_out_scores, d_out_pids,
        d_out_keys);
    CBK_CUDA(cudaGetLastError());
    count_launch();
    return CBK_OK;
  }
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
                        float* out_scores, int64_t* out_pids, uint64_t* out_keys, cudaStream_t stream,
                        const SelState* gate = nullptr) {
  const int n_groups = (n_lists + G - 1) / G;
  if (n_groups == 1 && !gate && static_cast<int64_t>(n_lists) * k_in <= 1024 && k <= 32) {
    merge_small_kernel<<<static_cast<unsigned int>(n_queries), kSmallThreads, 0, stream>>>(in, n_lists, sw, sq, k_in, k, out_scores,
                                                                                            out_pids, out_keys);
    CBK_CUDA(cudaGetLastError());
    count_launch();
    return CBK_OK;
  }
  const int P = padded_pow2(static_cast<int64_t>(std::min(G, n_lists)) * k_in);
  const size_t smem = static_cast<size_t>(P) * sizeof(uint64_t);
  const int threads = P >= 4096 ? 1024 : kTopkThreads;
  CBK_CUDA(cudaFuncSetAttribute(merge_topk_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  merge_topk_keys_kernel<<<dim3(n_groups, static_cast<unsigned int>(n_queries)), threads, smem, stream>>>(
      in, n_lists, G, sw, sq, k_in, P, k, out_scores, out_pids, out_keys, gate);
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
constexpr int64_t kRadixMinDocs = 2 * kDenseChunk;      // shorter rows: one or two chunk sorts are cheaper

static size_t chunk_sort_workspace_bytes(int64_t n_queries, int64_t n_docs, int k) {
  const int64_t nch = (n_docs + kDenseChunk - 1) / kDenseChunk;
  const int64_t k1 = std::min<int64_t>(k, kDenseChunk);
  const int64_t G = std::max<int64_t>(2, kTopkMaxCand / k1);
  const int64_t lvl2 = (nch + G - 1) / G;
  return static_cast<size_t>(n_queries * (nch + lvl2) * k1) * sizeof(uint64_t) + 256;
}

// radix-select scratch: [SelState × B | 3 histograms × B × kSelBins | candidates B × kTopkMaxCand keys]
static size_t radix_zeroed_bytes(int64_t n_queries) {
  return static_cast<size_t>(n_queries) * (sizeof(SelState) + 3 * kSelBins * sizeof(uint32_t));
}
static size_t radix_workspace_bytes(int64_t n_queries) {
  return ((radix_zeroed_bytes(n_queries) + 255) & ~static_cast<size_t>(255)) +
         static_cast<size_t>(n_queries) * kTopkMaxCand * sizeof(uint64_t);
}

size_t topk_dense_workspace_bytes(int64_t n_queries, int64_t n_docs, int k) {
  size_t b = chunk_sort_workspace_bytes(n_queries, n_docs, k);
  if (n_docs > kRadixMinDocs) b += radix_workspace_bytes(n_queries) + 256;
  return b;
}

// the chunk-sort path: every 16384-wide chunk is sorted, the per-chunk winners are merged level by level
static int chunk_sort_dispatch(const float* d_scores, int64_t n_queries, int64_t n_docs, int k, int64_t pid_base, int as_keys,
                               float* d_out_scores, int64_t* d_out_pids, void* d_workspace, const SelState* gate,
                               cudaStream_t stream) {
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
      d_scores, n_docs, kDenseChunk, pid_base, k1, bufA, gate);
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
                          last ? (as_keys ? reinterpret_cast<uint64_t*>(d_out_pids) : nullptr) : out, stream, gate);
    if (rc != CBK_OK) return rc;
    if (last) break;
    n_lists = n_groups;
    std::swap(in, out);
  }
  return CBK_OK;
}

int topk_dense_dispatch(const float* d_scores, int64_t n_queries, int64_t n_docs, int k, int64_t pid_base, int as_keys,
                        float* d_out_scores, int64_t* d_out_pids, void* d_workspace, cudaStream_t stream) {
  if (n_docs <= kRadixMinDocs)
    return chunk_sort_dispatch(d_scores, n_queries, n_docs, k, pid_base, as_keys, d_out_scores, d_out_pids, d_workspace, nullptr,
                               stream);
  // radix select; the chunk-sort kernels follow, gated per query on SelState::overflow (more than kTopkMaxCand scores
  // tied with the k-th — they exit at once otherwise), so the result is exact without a host round trip
  char* ws = static_cast<char*>(d_workspace);
  ws += (256 - (reinterpret_cast<uintptr_t>(ws) & 255)) & 255;
  SelState* state = reinterpret_cast<SelState*>(ws);
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + static_cast<size_t>(n_queries) * sizeof(SelState));
  uint64_t* cand = reinterpret_cast<uint64_t*>(ws + ((radix_zeroed_bytes(n_queries) + 255) & ~static_cast<size_t>(255)));
  void* chunk_ws = ws + radix_workspace_bytes(n_queries);
  CBK_CUDA(cudaMemsetAsync(state, 0, radix_zeroed_bytes(n_queries), stream));
  // scores per CTA: enough CTAs for four waves over the SMs, between 2048 and kSelChunk scores each
  int64_t want = (n_docs * n_queries + 4ll * sm_count() - 1) / (4ll * sm_count());
  want = ((want + kSelThreads - 1) / kSelThreads) * kSelThreads;
  const int chunk = static_cast<int>(std::min<int64_t>(kSelChunk, std::max<int64_t>(2048, want)));
  const dim3 grid(static_cast<unsigned int>((n_docs + chunk - 1) / chunk), static_cast<unsigned int>(n_queries));
  const unsigned int nq = static_cast<unsigned int>(n_queries);
  const size_t hstride = static_cast<size_t>(n_queries) * kSelBins;
  radix_hist_kernel<0><<<grid, kSelThreads, 0, stream>>>(d_scores, n_docs, chunk, state, hist);
  radix_select_kernel<0><<<nq, kSelThreads, 0, stream>>>(hist, state, k);
  radix_hist_kernel<1><<<grid, kSelThreads, 0, stream>>>(d_scores, n_docs, chunk, state, hist + hstride);
  radix_select_kernel<1><<<nq, kSelThreads, 0, stream>>>(hist + hstride, state, k);
  radix_hist_kernel<2><<<grid, kSelThreads, 0, stream>>>(d_scores, n_docs, chunk, state, hist + 2 * hstride);
  radix_select_kernel<2><<<nq, kSelThreads, 0, stream>>>(hist + 2 * hstride, state, k);
  radix_compact_kernel<<<grid, kSelThreads, 0, stream>>>(d_scores, n_docs, chunk, pid_base, state, cand);
  const size_t smem = static_cast<size_t>(kTopkMaxCand) * sizeof(uint64_t);
  CBK_CUDA(cudaFuncSetAttribute(radix_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  radix_final_kernel<<<nq, 1024, smem, stream>>>(cand, state, k, as_keys ? nullptr : d_out_scores, as_keys ? nullptr : d_out_pids,
                                                 as_keys ? reinterpret_cast<uint64_t*>(d_out_pids) : nullptr);
  CBK_CUDA(cudaGetLastError());
  count_launch(8);
  return chunk_sort_dispatch(d_scores, n_queries, n_docs, k, pid_base, as_keys, d_out_scores, d_out_pids, chunk_ws, state, stream);
}

}  // namespace cbk
