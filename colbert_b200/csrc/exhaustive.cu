// Query-batched exhaustive MaxSim (sm_100a, tcgen05 + TMEM + TMA): every document of the store (or
// of this GPU's shard) is scored against a batch of queries; the document tile is read from HBM once
// per pass of up to 16 queries and reused for all of them out of shared memory.
//
// This is the all-pairs shape of BaseModel.score (reference colbert/modeling/BaseModel.py:39-46,
// "qmh,dnh->qdmn" → max over n → sum over m) applied to the flat index store of
// colbert/ranking/colbert_ranker.py:61-73 — SURVEY.md §8d configs 4 and 5.
//
// Orientation: S[query rows, tokens] = Qblock[128, 128] · Dtile[128 tokens, 128]^T
//   * A operand = a block of 4 queries × 32 rows (fp32 → 16-bit, packed once per call, resident in
//     shared memory for a whole pass);  B operand = 128 consecutive store rows (one TMA box pair);
//   * accumulator = 128 TMEM lanes (query rows) × 128 columns (tokens): the max over a document's
//     tokens is a per-thread running max over columns (no cross-lane traffic), the sum over a query's
//     32 rows is one warp reduction per (query, document);
//   * document boundaries come from a bitmap with one bit per store row (set on the last row of each
//     document), so the epilogue never chases pfxsum;
//   * warp roles: warp 0 TMA producer, warps 1 and 18 MMA issuers (one elected lane each), warps 2-17 epilogue: four
//     groups of four warps (one TMEM lane quadrant = one query each).  A CTA owns four document-aligned token
//     sub-ranges and interleaves their tiles; group g drains the accumulators of sub-range g, so four
//     accumulators are being reduced while the next ones are being multiplied.  3-stage smem ring for
//     document tiles, 4 TMEM accumulator slots of 128 columns.
// bf16 stores: both MMA operands must share a format (mixed fp16 × bf16 is an illegal instruction),
// so the query is split into bf16 hi + lo parts and each tile is multiplied twice into the same
// accumulator (K = 256); fp16 stores need one pass over K.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "umma.cuh"

namespace cbk {

namespace {

constexpr int kTileTok = 128;
constexpr int kSmemTiles = 6;              // 32 KB tiles in shared memory: A blocks + B stages
constexpr int kMaxBStages = 5;
constexpr int kPend = 8;                  // per-warp staging of finished documents before a transposed sum
constexpr int kABlocks = 4;               // 32 KB A slots in shared memory
constexpr int kAccSlots = 4;              // × 128 TMEM columns
constexpr int kTileBytes = kTileTok * 256;
constexpr int kEpiGroups = 4;             // epilogue warp groups (4 warps each, one per TMEM lane quadrant)
constexpr int kIssuers = 2;               // MMA-issuing warps: one warp sustains one 128x128x16 MMA per ~90 cycles, two reach the 64-cycle floor
// Warp roles, by warpgroup: warps 0-15 = four epilogue groups (group = warp / 4, TMEM lane quadrant = warp % 4),
// warp 16 = TMA producer, warps 17-18 = MMA issuers, warp 19 idle.
constexpr int kEpiWarps = kEpiGroups * 4;
constexpr int kProducerWarp = kEpiWarps;
constexpr int kIssuer0Warp = kEpiWarps + 1;        // issues for groups 0-1
constexpr int kIssuer1Warp = kEpiWarps + 2;        // issues for groups 2-3
constexpr int kExhThreads = (kEpiWarps + 4) * 32;
constexpr int kEndsRing = 8;              // tiles PER SUB-RANGE whose document-end words are staged in shared memory (producer → epilogue)


struct StrideSet {
  int n;
  int v[CBK_MAX_STRIDES];
};

struct ExhMaps {
  CUtensorMap store;  // [rows, 128] 16-bit, box {64, 128}
  CUtensorMap q;      // packed queries [parts * n_qblocks * 128, 128] 16-bit, box {64, 128}
};

// ---- index-time metadata: bit t set ⇔ store row t is the last row of a document -------------------
__global__ void doc_end_bits_kernel(const int64_t* __restrict__ pfxsum, int64_t n_docs, uint32_t* __restrict__ bits) {
  const int64_t d = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (d >= n_docs) return;
  const int64_t end = pfxsum[d + 1] - 1;
  if (end >= pfxsum[d]) atomicOr(&bits[end >> 5], 1u << (end & 31));
}

// ---- per-launch: document-aligned, token-balanced ranges for the CTAs -------------------------------
__global__ void plan_ranges_kernel(const int64_t* __restrict__ pfxsum, int64_t n_docs, int n_ctas,
                                   int64_t* __restrict__ cta_doc_start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_ctas) return;
  const int64_t total = pfxsum[n_docs];
  const int64_t target = (total / n_ctas) * i + ((total % n_ctas) * i) / n_ctas;
  int64_t lo = 0, hi = n_docs;  // first doc d with pfxsum[d] >= target
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (pfxsum[mid] < target) lo = mid + 1; else hi = mid;
  }
  cta_doc_start[i] = i == n_ctas ? n_docs : lo;
}

// ---- per-launch: fp32 queries → 16-bit blocks of 4 queries × 32 rows, zero padded -------------------
// out rows [part][qblock][query-in-block][row] ; part 0 = value rounded to T (hi), part 1 = residual (lo)
template <typename T>
__global__ void pack_queries_kernel(const float* __restrict__ Q, int n_queries, int q_len, int n_qblocks, int parts,
                                    T* __restrict__ out) {
  const int64_t n_rows = static_cast<int64_t>(n_qblocks) * 128;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // one thread per (row, 4 columns)
  if (idx >= n_rows * 32) return;
  const int64_t row = idx >> 5;
  const int c = static_cast<int>(idx & 31) * 4;
  const int q = static_cast<int>(row >> 5), r = static_cast<int>(row & 31);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q < n_queries && r < q_len) v = *reinterpret_cast<const float4*>(Q + (static_cast<int64_t>(q) * q_len + r) * 128 + c);
  const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const T hi = static_cast<T>(in[j]);
    out[row * 128 + c + j] = hi;
    if (parts > 1) out[(n_rows + row) * 128 + c + j] = static_cast<T>(in[j] - static_cast<float>(hi));
  }
}

// Per-warp staging of finished documents: at a document end every lane (= query row) parks its row
// maximum in pend.v[slot][lane] (and the document length in pend.len[slot]) — a handful of instructions at
// each of the 32 unrolled column positions, which keeps the epilogue inside the instruction cache.  When
// kPend documents are parked, lane d applies the reference's zero floor (doclen ∉ strides, SURVEY.md §8
// a12') and adds up the 32 rows of document d (a transposed sum: no shuffle chain per document), and the
// warp stores kPend consecutive scores at once.
struct __align__(16) PendBuf {
  float v[kPend][33];
  int len[kPend];
  int n_strides;
  int strides[CBK_MAX_STRIDES];
};

__device__ __noinline__ void flush_pending(PendBuf* pb, int n, float* __restrict__ dst_first, bool write) {
  __syncwarp();
  const int lane = threadIdx.x & 31;
  static_assert(kPend == 8, "flush_pending maps lane -> (document = lane / 4, quarter of the query rows = lane % 4)");
  const int d = lane >> 2, qtr = lane & 3;
  const int doclen = pb->len[d];                      // (documents >= n: stale values, computed and dropped)
  const int n_strides = pb->n_strides;
  float floor_v = n_strides > 0 ? 0.f : -INFINITY;
  for (int i = 0; i < n_strides; ++i)
    if (pb->strides[i] == doclen) floor_v = -INFINITY;
  const float* row = pb->v[d] + qtr * 8;              // bank (33 d + 8 qtr + i) % 32: conflict-free across the warp
  float s0 = fmaxf(row[0], floor_v) + fmaxf(row[1], floor_v);
  float s1 = fmaxf(row[2], floor_v) + fmaxf(row[3], floor_v);
  float s2 = fmaxf(row[4], floor_v) + fmaxf(row[5], floor_v);
  float s3 = fmaxf(row[6], floor_v) + fmaxf(row[7], floor_v);
  float sum = (s0 + s1) + (s2 + s3);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  if (write && qtr == 0 && d < n) dst_first[d] = sum;
  __syncwarp();
}

// Epilogue state of one warp: st.r is each lane's running maximum over the open document's rows seen so far;
// doc_start (row index, relative to the warp's sub-range, of the open document's first row) and docs_done
// (documents closed so far in this pass; docs_done % kPend of them are parked in the PendBuf) are warp-uniform.
struct EpiState {
  float r;
  int doc_start;
  int docs_done;
};

// The 8 accumulator columns x0..x7 (rows col..col+7 of the sub-range) hold at least one document end (m8: bit j
// set = column j is the last row of a document).  One out-of-line copy serves every call site, which keeps the
// epilogue loop itself small enough to stay in the instruction cache.
__device__ __noinline__ EpiState close_docs_in_group(PendBuf* pb, EpiState st, float x0, float x1, float x2, float x3, float x4,
                                                     float x5, float x6, float x7, uint32_t m8, int col, float* dst_row,
                                                     int write) {
  const int lane = threadIdx.x & 31;
  const float x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    st.r = fmaxf(st.r, x[j]);
    if ((m8 >> j) & 1u) {   // warp-uniform
      const int slot = st.docs_done & (kPend - 1);
      pb->v[slot][lane] = st.r;
      if (lane == 0) pb->len[slot] = col + j + 1 - st.doc_start;
      st.doc_start = col + j + 1;
      ++st.docs_done;
      st.r = -INFINITY;
      if ((st.docs_done & (kPend - 1)) == 0) flush_pending(pb, kPend, dst_row + (st.docs_done - kPend), write != 0);
    }
  }
  return st;
}

// max of 8 consecutive accumulator columns and the running value
__device__ __forceinline__ float max8(const uint32_t* v, float r) {
  const float x0 = fmaxf(fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), __uint_as_float(v[2]));
  const float x1 = fmaxf(fmaxf(__uint_as_float(v[3]), __uint_as_float(v[4])), __uint_as_float(v[5]));
  const float x2 = fmaxf(fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7])), r);
  return fmaxf(fmaxf(x0, x1), x2);
}

// max of accumulator columns v[L .. R-1] as a ternary tree (one FMNMX3 per node)
template <int L, int R>
__device__ __forceinline__ float max_range(const uint32_t (&v)[32]) {
  constexpr int n = R - L;
  static_assert(n >= 1, "empty range");
  if constexpr (n == 1) {
    return __uint_as_float(v[L]);
  } else if constexpr (n == 2) {
    return fmaxf(__uint_as_float(v[L]), __uint_as_float(v[L + 1]));
  } else {
    constexpr int a = L + (n + 2) / 3, b = a + (R - a + 1) / 2;
    return fmaxf(fmaxf(max_range<L, a>(v), max_range<a, b>(v)), max_range<b, R>(v));
  }
}

// Exactly ONE document ends inside the 32 columns, at column E: head = max(r, v[0..E]) closes it, tail = max(v[E+1..31])
// opens the next one.  One straight-line copy per E (16-17 three-input maxima each), selected by a jump table: a chunk
// with a document end costs about as many instructions as one without.
template <int E>
__device__ __forceinline__ void split32(const uint32_t (&v)[32], float r, float& head, float& tail) {
  head = fmaxf(r, max_range<0, E + 1>(v));
  if constexpr (E < 31) tail = max_range<E + 1, 32>(v);
  else tail = -INFINITY;
}

__device__ __forceinline__ void split32_at(const uint32_t (&v)[32], int e, float r, float& head, float& tail) {
#define CBK_S32(E) case E: split32<E>(v, r, head, tail); break;
  switch (e) {
    CBK_S32(0) CBK_S32(1) CBK_S32(2) CBK_S32(3) CBK_S32(4) CBK_S32(5) CBK_S32(6) CBK_S32(7)
    CBK_S32(8) CBK_S32(9) CBK_S32(10) CBK_S32(11) CBK_S32(12) CBK_S32(13) CBK_S32(14) CBK_S32(15)
    CBK_S32(16) CBK_S32(17) CBK_S32(18) CBK_S32(19) CBK_S32(20) CBK_S32(21) CBK_S32(22) CBK_S32(23)
    CBK_S32(24) CBK_S32(25) CBK_S32(26) CBK_S32(27) CBK_S32(28) CBK_S32(29) CBK_S32(30)
    default: split32<31>(v, r, head, tail); break;
  }
#undef CBK_S32
}

// One 32-column chunk of an accumulator: fold the columns into the running maximum of the open document, closing
// every document whose last row lies inside the chunk (m: bit j = column j ends a document; col = row index of
// column 0 relative to the warp's sub-range).
__device__ __forceinline__ void drain_chunk(const uint32_t (&v)[32], uint32_t m, int col, PendBuf* pb, EpiState& st,
                                            float* dst_row, int write, int lane) {
  if (m == 0u) {   // no document ends inside these 32 columns (the common case): one max tree
    const float y0 = max8(v, st.r), y1 = max8(v + 8, -INFINITY), y2 = max8(v + 16, -INFINITY), y3 = max8(v + 24, -INFINITY);
    st.r = fmaxf(fmaxf(y0, y1), fmaxf(y2, y3));
    return;
  }
  if ((m & (m - 1u)) == 0u) {   // one document ends in these 32 columns (any index whose documents have 32+ rows)
    const int e = 31 - __clz(static_cast<int>(m));
    float head, tail;
    split32_at(v, e, st.r, head, tail);
    const int slot = st.docs_done & (kPend - 1);
    const int end1 = col + e + 1;
    pb->v[slot][lane] = head;
    if (lane == 0) pb->len[slot] = end1 - st.doc_start;
    st.doc_start = end1;
    ++st.docs_done;
    st.r = tail;
    if ((st.docs_done & (kPend - 1)) == 0) flush_pending(pb, kPend, dst_row + (st.docs_done - kPend), write != 0);
    return;
  }
  // several documents end inside the chunk (documents shorter than 32 rows): 8 columns at a time, out of line
#pragma unroll
  for (int s8 = 0; s8 < 4; ++s8) {
    const uint32_t m8 = (m >> (8 * s8)) & 0xffu;
    if (m8 == 0u) {
      st.r = max8(v + 8 * s8, st.r);
    } else {
      const uint32_t* x = v + 8 * s8;
      st = close_docs_in_group(pb, st, __uint_as_float(x[0]), __uint_as_float(x[1]), __uint_as_float(x[2]),
                               __uint_as_float(x[3]), __uint_as_float(x[4]), __uint_as_float(x[5]),
                               __uint_as_float(x[6]), __uint_as_float(x[7]), m8, col + 8 * s8, dst_row, write);
    }
  }
}

// =====================================================================================================
// kStats: per-role cycle accounting (debug builds of the launch: CBK_EXH_STATS=1), four counters per warp:
//   producer  [0] waiting for a free stage                                              [3] total
//   issuers   [0] waiting for a full stage   [1] waiting for a free accumulator         [3] total
//   epilogue  [0] waiting for a full accumulator   [1] draining it   [2] tile prologue   [3] total
template <bool kStats>
__global__ void __launch_bounds__(kExhThreads, 1)
maxsim_exhaustive_kernel(const __grid_constant__ ExhMaps maps, const uint32_t* __restrict__ doc_end_bits,
                         const int64_t* __restrict__ pfxsum, const int64_t* __restrict__ cta_doc_start,
                         StrideSet strides, int n_queries, int n_qblocks, int parts, int64_t n_docs, uint32_t idesc,
                         float* __restrict__ scores, long long* __restrict__ stats) {
  long long sc0 = 0, sc1 = 0, sc2 = 0;
  const long long sc_begin = kStats ? clock64() : 0;
#define CBK_T0() long long _t0 = kStats ? clock64() : 0
#define CBK_T1(acc) if (kStats) acc += clock64() - _t0
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a_full, bar_pass_done;
  __shared__ __align__(8) uint64_t bar_b_full[kMaxBStages], bar_b_empty[kMaxBStages];
  // acc_full is per (consuming epilogue group, slot): an mbarrier wait only tells phases apart by parity, so a
  // barrier must never have two waiters that are a whole phase apart (two groups sharing one slot would be)
  __shared__ __align__(8) uint64_t bar_acc_full[kEpiGroups], bar_acc_empty[kEpiGroups];   // group g owns TMEM slot g
  __shared__ uint32_t tmem_base_smem;
  __shared__ PendBuf s_pend[kEpiGroups * 4];
  // document-end bits of the tiles in flight, one uint4 per tile: word c bit j = column 32c + j is the last row of a
  // document OF THE TILE'S SUB-RANGE.  Written by the producer warp before it arms the tile's stage barrier (so the
  // words are visible to whoever observes, through the MMA, that the tile has arrived), read by the 16 epilogue warps:
  // what used to be 5 global loads, 4 funnel shifts and the last-tile masking in every epilogue warp.
  // Indexed [sub-range][tile % kEndsRing]: the producer is never more than 1 + kMaxBStages tiles of a sub-range ahead of
  // the group that drains it (tile t + 1 cannot be multiplied before the group released tile t, and it read the words
  // of tile t before that), so 8 entries per sub-range are never overwritten while still unread.
  __shared__ __align__(16) uint32_t s_ends[kEpiGroups][kEndsRing][4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp < kEpiWarps && lane <= CBK_MAX_STRIDES) {
    if (lane == 0) s_pend[warp].n_strides = strides.n;
    else s_pend[warp].strides[lane - 1] = strides.v[lane - 1];
  }
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_addr = (raw + 1023u) & ~1023u;              // kABlocks × 32 KB
  // the A region holds as many 32 KB blocks as the largest pass needs; the rest of the tiles are B stages
  const int a_tiles = min(kABlocks, n_qblocks * parts);
  const uint32_t b_addr = a_addr + a_tiles * kTileBytes;
  const uint32_t kBStages = min(kMaxBStages, kSmemTiles - a_tiles);

  if (tid == 0) {
    mbar_init(smem_u32(&bar_a_full), 1);
    mbar_init(smem_u32(&bar_pass_done), kIssuers);   // every issuer commits once per pass
    for (int s = 0; s < kMaxBStages; ++s) {
      mbar_init(smem_u32(&bar_b_full[s]), 1);
      mbar_init(smem_u32(&bar_b_empty[s]), kIssuers);   // every issuer commits once per tile
    }
    for (int s = 0; s < kEpiGroups; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), 4);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == kIssuer0Warp) {
    umma::tmem_alloc(smem_u32(&tmem_base_smem), 512);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;

  // A CTA owns 4 consecutive document-aligned token ranges.  In a pass with qb query blocks they are merged
  // into n_sub = 4 / qb_eff sub-ranges (qb_eff = qb rounded up to 1, 2 or 4) whose tiles are interleaved;
  // item (t, s) = tile t of sub-range s carries qb accumulators, and epilogue group g = s*qb_eff + a drains
  // accumulator a of sub-range s — so every group follows ONE query block over ONE contiguous run of documents.
  int64_t rng_d[kEpiGroups + 1], rng_tok[kEpiGroups + 1];
#pragma unroll
  for (int g = 0; g <= kEpiGroups; ++g) {
    rng_d[g] = cta_doc_start[blockIdx.x * kEpiGroups + g];
    rng_tok[g] = pfxsum[rng_d[g]];
  }
  const int qb_max = kABlocks / parts;                          // query blocks per pass
  const int n_passes = (n_qblocks + qb_max - 1) / qb_max;

  // tiles of sub-range s when the 4 ranges are merged into n_sub sub-ranges
  auto sub_tiles = [&](int n_sub, int s) -> int {
    const int w = kEpiGroups / n_sub;
    int64_t t0 = 0, t1 = 0;
#pragma unroll
    for (int g = 0; g <= kEpiGroups; ++g) {
      if (g == s * w) t0 = rng_tok[g];
      if (g == (s + 1) * w) t1 = rng_tok[g];
    }
    return static_cast<int>((t1 - t0 + kTileTok - 1) / kTileTok);
  };
  auto sub_tok0 = [&](int n_sub, int s) -> int64_t {
    const int w = kEpiGroups / n_sub;
    int64_t t0 = 0;
#pragma unroll
    for (int g = 0; g <= kEpiGroups; ++g)
      if (g == s * w) t0 = rng_tok[g];
    return t0;
  };

  if (warp >= kEpiWarps) {
   if (warp == kProducerWarp) {
    // ===================================== TMA producer =============================================
    // (whole warp, uniform control flow; one elected lane issues — see elect_one in cbk_common.cuh)
    {
      if (lane == 0) {
        tma_prefetch_desc(&maps.store);
        tma_prefetch_desc(&maps.q);
      }
      uint32_t it = 0;
      uint32_t ring0[kEpiGroups] = {0u, 0u, 0u, 0u};   // tiles of sub-range index s2 in the passes before this one (s_ends position)
      for (int p = 0; p < n_passes; ++p) {
        if (p > 0) mbar_wait(smem_u32(&bar_pass_done), (p - 1) & 1);   // every MMA that read the old A blocks is done
        const int qb = min(qb_max, n_qblocks - p * qb_max);
        const int n_sub = qb == 1 ? 4 : (qb == 2 ? 2 : 1);
        const uint32_t afull = smem_u32(&bar_a_full);
        if (elect_one()) {
          mbar_arrive_expect_tx(afull, static_cast<uint32_t>(qb * parts) * kTileBytes);
          for (int a = 0; a < qb; ++a)
            for (int part = 0; part < parts; ++part) {
              const int row = (part * n_qblocks + p * qb_max + a) * 128;
              const uint32_t dst = a_addr + (a * parts + part) * kTileBytes;
              tma_load_2d(dst, &maps.q, 0, row, afull, kEvictLast);
              tma_load_2d(dst + kTileBytes / 2, &maps.q, 64, row, afull, kEvictLast);
            }
        }
        __syncwarp();
        int nt[kEpiGroups], rows[kEpiGroups];
        int64_t t0[kEpiGroups];
        int max_nt = 0;
#pragma unroll
        for (int s2 = 0; s2 < kEpiGroups; ++s2) {
          nt[s2] = s2 < n_sub ? sub_tiles(n_sub, s2) : 0;
          t0[s2] = s2 < n_sub ? sub_tok0(n_sub, s2) : 0;
          rows[s2] = s2 < n_sub ? static_cast<int>((s2 + 1 < n_sub ? sub_tok0(n_sub, s2 + 1) : rng_tok[kEpiGroups]) - t0[s2]) : 0;
          max_nt = max(max_nt, nt[s2]);
        }
        for (int t = 0; t < max_nt; ++t)
#pragma unroll
          for (int s2 = 0; s2 < kEpiGroups; ++s2) {
            if (t >= nt[s2]) continue;
            const uint32_t st = it % kBStages;
            // document-end words of this tile (lanes 0-4 fetch one bitmap word each; they arrive while the warp waits
            // for the stage): bit j of word c = column 32c + j, ends past the sub-range's last row belong to a neighbour
            const uint32_t bw = lane < 5 ? doc_end_bits[(t0[s2] >> 5) + 4 * t + lane] : 0u;
            { CBK_T0(); mbar_wait(smem_u32(&bar_b_empty[st]), ((it / kBStages) & 1u) ^ 1u); CBK_T1(sc0); }
            {
              uint32_t e = __funnelshift_r(bw, __shfl_down_sync(0xffffffffu, bw, 1), static_cast<uint32_t>(t0[s2]) & 31u);
              const int l = rows[s2] - t * kTileTok - 32 * lane;
              if (l <= 0) e = 0u;
              else if (l < 32) e &= (1u << l) - 1u;
              if (lane < 4) s_ends[s2][(ring0[s2] + t) % kEndsRing][lane] = e;
              __syncwarp();
            }
            const uint32_t full = smem_u32(&bar_b_full[st]);
            const uint32_t dst = b_addr + st * kTileBytes;
            const int row = static_cast<int>(t0[s2]) + t * kTileTok;
            if (elect_one()) {
              mbar_arrive_expect_tx(full, kTileBytes);
              tma_load_2d(dst, &maps.store, 0, row, full, kEvictFirst);
              tma_load_2d(dst + kTileBytes / 2, &maps.store, 64, row, full, kEvictFirst);
            }
            __syncwarp();
            ++it;
          }
#pragma unroll
        for (int s2 = 0; s2 < kEpiGroups; ++s2) ring0[s2] += static_cast<uint32_t>(nt[s2]);
      }
    }
  } else if (warp == kIssuer0Warp || warp == kIssuer1Warp) {
    // ===================================== MMA issuers ==============================================
    // Two warps, each with uniform control flow and one elected lane issuing (see elect_one in cbk_common.cuh):
    // issuer w feeds epilogue groups 2w and 2w+1.  A single issuing warp sustains one 128x128x16 MMA per ~90
    // cycles, two reach the tensor pipe's 64-cycle floor (benchmarks/umma_rate.py).  These warps pace the kernel,
    // so their loop is kept short: barrier addresses and descriptor words are formed once, ring positions advance
    // by increments, and there is one copy of the issue code.  Both issuers walk every tile (wait for it, commit
    // its release) even when it carries no accumulator of theirs, which keeps the stage barriers' counts fixed.
    const int w = warp - kIssuer0Warp;
    const uint32_t bfull0 = hold(smem_u32(&bar_b_full[0])), bempty0 = hold(smem_u32(&bar_b_empty[0]));
    const uint32_t accfull0 = hold(smem_u32(&bar_acc_full[0])), accempty0 = hold(smem_u32(&bar_acc_empty[0]));
    const uint32_t a_lo0 = hold(umma::desc_lo_sw128(a_addr)), b_lo0 = hold(umma::desc_lo_sw128(b_addr));
    constexpr uint32_t kTileDesc = kTileBytes >> 4, kHalfDesc = (kTileBytes / 2) >> 4;
    uint32_t st = 0, st_parity = 0;                       // B stage ring position / parity of its next full phase
    uint32_t empty_parity = 0;                            // bit g = parity of the next wait on bar_acc_empty[g]
    for (int p = 0; p < n_passes; ++p) {
      const int qb = min(qb_max, n_qblocks - p * qb_max);
      const int n_sub = qb == 1 ? 4 : (qb == 2 ? 2 : 1);
      const int qb_eff = kEpiGroups / n_sub;
      const int nt0 = sub_tiles(n_sub, 0), nt1 = n_sub > 1 ? sub_tiles(n_sub, 1) : 0;
      const int nt2 = n_sub > 2 ? sub_tiles(n_sub, 2) : 0, nt3 = n_sub > 3 ? sub_tiles(n_sub, 3) : 0;
      const int max_nt = max(max(nt0, nt1), max(nt2, nt3));
      mbar_wait(smem_u32(&bar_a_full), p & 1);
      umma::fence_after_sync();
      for (int t = 0; t < max_nt; ++t) {
#pragma unroll 1
        for (int s2 = 0; s2 < n_sub; ++s2) {
          const int nts = s2 == 0 ? nt0 : (s2 == 1 ? nt1 : (s2 == 2 ? nt2 : nt3));
          if (t >= nts) continue;
          { CBK_T0(); mbar_wait(bfull0 + 8 * st, st_parity); CBK_T1(sc0); }
          umma::fence_after_sync();
          const uint32_t b_lo = b_lo0 + st * kTileDesc;
#pragma unroll 1
          for (int a = 0; a < qb; ++a) {
            const uint32_t g = static_cast<uint32_t>(s2 * qb_eff + a);     // epilogue group = TMEM slot
            if (static_cast<int>(g >> 1) != w) continue;
            { CBK_T0(); mbar_wait(accempty0 + 8 * g, ((empty_parity >> g) & 1u) ^ 1u); CBK_T1(sc1); }
            empty_parity ^= 1u << g;
            umma::fence_after_sync();
            const uint32_t d_tmem = tmem + g * kTileTok;
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(a * parts) * kTileDesc;
            if (elect_one()) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma::mma_f16_ss_lo(d_tmem, a_lo + h * kHalfDesc + 2 * k, b_lo + h * kHalfDesc + 2 * k, idesc, (h | k) ? 1u : 0u);
              if (parts > 1) {   // bf16 store: the lo part of the query, same accumulator
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma::mma_f16_ss_lo(d_tmem, a_lo + kTileDesc + h * kHalfDesc + 2 * k, b_lo + h * kHalfDesc + 2 * k, idesc, 1u);
              }
              umma::commit(accfull0 + 8 * g);
            }
            __syncwarp();
          }
          if (elect_one()) umma::commit(bempty0 + 8 * st);
          __syncwarp();
          if (++st == kBStages) {
            st = 0;
            st_parity ^= 1u;
          }
        }
      }
      if (elect_one()) umma::commit(smem_u32(&bar_pass_done));
      __syncwarp();
    }
   }
  } else {
    // ===================================== epilogue (warps 0..15) ===================================
    const int grp = warp >> 2;                       // epilogue group
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read = query within the block
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    PendBuf* const pb = &s_pend[warp];
    uint32_t full_parity = 0;                        // parity of this group's next wait on its accumulator-full barrier
    uint32_t ring0[kEpiGroups] = {0u, 0u, 0u, 0u};   // as in the producer: tiles of sub-range index s2 in earlier passes
    for (int p = 0; p < n_passes; ++p) {
      const int qb = min(qb_max, n_qblocks - p * qb_max);
      const int n_sub = qb == 1 ? 4 : (qb == 2 ? 2 : 1);
      const int qb_eff = kEpiGroups / n_sub;
      const int my_s = grp / qb_eff, my_a = grp % qb_eff;
      const bool active = my_a < qb;
      int nt[kEpiGroups];
      int max_nt = 0;
#pragma unroll
      for (int s2 = 0; s2 < kEpiGroups; ++s2) {
        nt[s2] = s2 < n_sub ? sub_tiles(n_sub, s2) : 0;
        max_nt = max(max_nt, nt[s2]);
      }
      // my contiguous run of documents / rows
      const int wdt = kEpiGroups / n_sub;
      int64_t my_d0 = 0, my_tok0 = 0, my_tok1 = 0;
#pragma unroll
      for (int g = 0; g <= kEpiGroups; ++g) {
        if (g == my_s * wdt) {
          my_d0 = rng_d[g];
          my_tok0 = rng_tok[g];
        }
        if (g == (my_s + 1) * wdt) my_tok1 = rng_tok[g];
      }
      const int q = (p * qb_max + my_a) * 4 + quad;
      const int write = active && q < n_queries;
      float* const dst_row = scores + static_cast<int64_t>(write ? q : 0) * n_docs + my_d0;   // score of my first document
      int my_nt = 0;
      uint32_t my_ring0 = 0;
#pragma unroll
      for (int s2 = 0; s2 < kEpiGroups; ++s2) {
        if (s2 == my_s) {
          my_nt = nt[s2];
          my_ring0 = ring0[s2];
        }
        ring0[s2] += static_cast<uint32_t>(nt[s2]);
      }

      EpiState st;
      st.r = -INFINITY;
      st.doc_start = 0;
      st.docs_done = 0;

      for (int t = 0; t < max_nt; ++t) {
        if (!active || t >= my_nt) continue;
        const long long t_tile = kStats ? clock64() : 0;

        const long long t_wait = kStats ? clock64() : 0;
        mbar_wait(smem_u32(&bar_acc_full[grp]), full_parity);      // group g owns TMEM slot g
        const long long t_drain = kStats ? clock64() : 0;
        if (kStats) { sc2 += t_wait - t_tile; sc0 += t_drain - t_wait; }
        full_parity ^= 1u;
        umma::fence_after_sync();
        const uint32_t t_addr = tmem + lane_base + grp * kTileTok;
        const int col0 = t * kTileTok;
        // One 32-column chunk in registers at a time.  Measured and NOT adopted (B200, Nq = 16, 1.1 M-document shard):
        // a second register buffer with the load of chunk c+1 in flight while chunk c is folded — 13.0 ms against 9.3 ms;
        // two 64-column loads per accumulator, the slot handed back after the second — 10.4 ms against 7.9 ms (both with
        // setmaxnreg moving registers from the producer / issuer warpgroup to the epilogue warpgroups so that the hot
        // loop does not spill: 128 x 40 + 512 x 104 registers is all a 640-thread CTA owns).  A warp with a tcgen05.ld
        // outstanding does not overlap it with its own arithmetic here, and the epilogue body doubles in size.
        // 128 document-end bits of this tile, staged by the producer: bit j of word c is column 32c + j
        const uint4 ends = *reinterpret_cast<const uint4*>(&s_ends[my_s][(my_ring0 + t) % kEndsRing][0]);
        uint32_t m = ends.x, m1 = ends.y, m2 = ends.z, m3 = ends.w;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          umma::tmem_ld_32x32(t_addr + c * 32, v);
          umma::tmem_ld_wait();
          if (c == 3) {   // all columns of the slot are in registers: hand it back to the MMA warp
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[grp]));
          }
          drain_chunk(v, m, col0 + c * 32, pb, st, dst_row, write, lane);
          m = m1;
          m1 = m2;
          m2 = m3;
        }
        if (kStats) sc1 += clock64() - t_drain;
      }
      const int n_pend = st.docs_done & (kPend - 1);
      if (n_pend > 0) flush_pending(pb, n_pend, dst_row + (st.docs_done - n_pend), write != 0);
    }
  }

  if (kStats && lane == 0) {
    long long* o = stats + (static_cast<int64_t>(blockIdx.x) * (kExhThreads / 32) + warp) * 4;
    o[0] = sc0; o[1] = sc1; o[2] = sc2; o[3] = clock64() - sc_begin;
  }
#undef CBK_T0
#undef CBK_T1
  umma::fence_before_sync();
  __syncthreads();
  if (warp == kIssuer0Warp) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

// Profiling aid (CBK_EXH_STATS=1): the instrumented instantiation, synchronous, per-role mean cycle counts to stderr.
static int exhaustive_launch_with_stats(const ExhMaps& maps, const uint32_t* bits, const int64_t* pfxsum, const int64_t* ranges,
                                        const StrideSet& ss, int n_queries, int n_qblocks, int parts, int64_t n_docs,
                                        uint32_t idesc, float* out, int n_ctas, size_t smem, cudaStream_t stream) {
  constexpr int kW = kExhThreads / 32;
  long long* d_stats = nullptr;
  const size_t bytes = static_cast<size_t>(n_ctas) * kW * 4 * sizeof(long long);
  CBK_CUDA(cudaMalloc(&d_stats, bytes));
  CBK_CUDA(cudaMemsetAsync(d_stats, 0, bytes, stream));
  CBK_CUDA(cudaFuncSetAttribute(maxsim_exhaustive_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  maxsim_exhaustive_kernel<true><<<n_ctas, kExhThreads, smem, stream>>>(maps, bits, pfxsum, ranges, ss, n_queries, n_qblocks, parts,
                                                                      n_docs, idesc, out, d_stats);
  CBK_CUDA(cudaGetLastError());
  CBK_CUDA(cudaStreamSynchronize(stream));
  long long* h = static_cast<long long*>(std::malloc(bytes));
  CBK_CUDA(cudaMemcpy(h, d_stats, bytes, cudaMemcpyDeviceToHost));
  cudaFree(d_stats);
  double acc[kW][4] = {};
  for (int c = 0; c < n_ctas; ++c)
    for (int w = 0; w < kW; ++w)
      for (int k = 0; k < 4; ++k) acc[w][k] += static_cast<double>(h[(static_cast<size_t>(c) * kW + w) * 4 + k]) / n_ctas;
  std::free(h);
  std::fprintf(stderr, "[exh stats] n_queries=%d qblocks=%d ctas=%d (mean cycles per CTA)\n", n_queries, n_qblocks, n_ctas);
  for (int w = 0; w < kW; ++w)
    std::fprintf(stderr, "[exh stats] warp %2d: c0 %.0f c1 %.0f c2 %.0f total %.0f\n", w, acc[w][0], acc[w][1], acc[w][2], acc[w][3]);
  count_launch(3);
  return CBK_OK;
}

size_t doc_end_bits_bytes(int64_t n_store_rows) {
  return static_cast<size_t>((n_store_rows + 31) / 32 + 8) * sizeof(uint32_t);   // +8 words: tiles read 5 words past their start
}

int doc_end_bits_dispatch(const int64_t* d_pfxsum, int64_t n_docs, int64_t n_store_rows, uint32_t* d_bits,
                          cudaStream_t stream) {
  CBK_CUDA(cudaMemsetAsync(d_bits, 0, doc_end_bits_bytes(n_store_rows), stream));
  const int threads = 256;
  const unsigned int blocks = static_cast<unsigned int>((n_docs + threads - 1) / threads);
  doc_end_bits_kernel<<<blocks, threads, 0, stream>>>(d_pfxsum, n_docs, d_bits);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

size_t exhaustive_workspace_bytes(int64_t n_queries) {
  const int64_t n_qblocks = (n_queries + 3) / 4;
  return 8192 /* CTA sub-ranges */ + static_cast<size_t>(2 * n_qblocks * 128 * 128 * 2) /* packed queries, ≤ 2 parts */;
}

int exhaustive_dispatch(const void* d_store, int store_dtype, int64_t n_store_rows, const int64_t* d_pfxsum,
                        const uint32_t* d_doc_end_bits, int64_t n_docs, const int32_t* strides, int n_strides,
                        const float* d_Q, int q_len, int64_t n_queries, float* d_out_scores, void* d_workspace,
                        int flags, cudaStream_t stream) {
  const int n_ctas = static_cast<int>(std::min<int64_t>(sm_count(), std::max<int64_t>(1, n_docs / kEpiGroups)));
  const int n_qblocks = static_cast<int>((n_queries + 3) / 4);
  const bool bf16 = store_dtype == CBK_BF16;
  const int parts = (bf16 && !(flags & CBK_FLAG_BF16_NATIVE_MMA)) ? 2 : 1;
  int64_t* d_ranges = static_cast<int64_t*>(d_workspace);
  void* d_qp = static_cast<uint8_t*>(d_workspace) + 8192;

  plan_ranges_kernel<<<(n_ctas * kEpiGroups + 256) / 256, 256, 0, stream>>>(d_pfxsum, n_docs, n_ctas * kEpiGroups, d_ranges);
  CBK_CUDA(cudaGetLastError());
  const int64_t pack_threads = static_cast<int64_t>(n_qblocks) * 128 * 32;
  const unsigned int pack_blocks = static_cast<unsigned int>((pack_threads + 255) / 256);
  if (bf16)
    pack_queries_kernel<__nv_bfloat16><<<pack_blocks, 256, 0, stream>>>(d_Q, static_cast<int>(n_queries), q_len, n_qblocks,
                                                                        parts, static_cast<__nv_bfloat16*>(d_qp));
  else
    pack_queries_kernel<__half><<<pack_blocks, 256, 0, stream>>>(d_Q, static_cast<int>(n_queries), q_len, n_qblocks, parts,
                                                                 static_cast<__half*>(d_qp));
  CBK_CUDA(cudaGetLastError());

  ExhMaps maps;
  int rc = make_store_tensor_map(&maps.store, d_store, n_store_rows, 128, 64, kTileTok);
  if (rc != CBK_OK) return rc;
  rc = make_store_tensor_map(&maps.q, d_qp, static_cast<int64_t>(parts) * n_qblocks * 128, 128, 64, 128);
  if (rc != CBK_OK) return rc;
  StrideSet ss;
  ss.n = n_strides;
  for (int i = 0; i < CBK_MAX_STRIDES; ++i) ss.v[i] = i < n_strides ? strides[i] : -1;
  const uint32_t fmt = bf16 ? umma::kFmtBF16 : umma::kFmtF16;
  const uint32_t idesc = umma::make_idesc(128, kTileTok, fmt, fmt);
  const size_t smem = 1024 + static_cast<size_t>(kSmemTiles) * kTileBytes;
  static const bool want_stats = std::getenv("CBK_EXH_STATS") != nullptr;     // profiling aid, never set in production
  if (want_stats) return exhaustive_launch_with_stats(maps, d_doc_end_bits, d_pfxsum, d_ranges, ss, static_cast<int>(n_queries),
                                                      n_qblocks, parts, n_docs, idesc, d_out_scores, n_ctas, smem, stream);
  CBK_CUDA(cudaFuncSetAttribute(maxsim_exhaustive_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  maxsim_exhaustive_kernel<false><<<n_ctas, kExhThreads, smem, stream>>>(maps, d_doc_end_bits, d_pfxsum, d_ranges, ss,
                                                                       static_cast<int>(n_queries), n_qblocks, parts, n_docs, idesc,
                                                                       d_out_scores, nullptr);
  CBK_CUDA(cudaGetLastError());
  count_launch(3);
  return CBK_OK;
}

}  // namespace cbk
