// MaxSim rerank for wide embeddings on tensor cores (sm_100a): dim a multiple of 64 up to 1024, other than the
// 128 that maxsim_rerank_kernel is specialised for.  The author's own configuration scores 768-wide embeddings
// (reference proj_conf/dense.yaml:8, no projection), for which a row is 1536 B and the query no longer fits a warp's
// registers.  Same contract as cbk_maxsim_rerank (reference colbert_ranker.py:88-126, BaseModel.py:39-46).
//
// Split along K instead of along candidates:
//   * a CTA (8 warps) works on ONE document tile of 16 rows × dim at a time, streamed by TMA through a CTA-wide ring
//     of stages (one 3-D box {64 columns, 16 rows, dim/64 slabs} per tile, 128-byte swizzle per slab);
//   * warp w owns the k-steps [ks0, ks0 + nks) of the dim/16 in a row and keeps ITS slice of the query as mma.sync
//     A fragments in registers (≤ 64 registers, as in the 128-wide kernel) — the query never sits in shared memory,
//     so shared-memory bandwidth is spent on document rows only (a token-split would re-read the whole query, 2 × the
//     document bytes, per tile);
//   * each warp multiplies its slice of the tile (ldmatrix.x4 + m16n8k16, fp32 accumulate) into a partial
//     S[32 query rows, 16 tokens] and parks it in shared memory; after one __syncthreads warp w' adds up the eight
//     partials of two accumulator registers (= 8 query rows × 8 tokens), masks the tokens past the document's end and
//     folds them into a running maximum.  Partials are double-buffered: one block barrier per tile;
//   * at a document's end the two half-tile maxima of each query row meet in shared memory, the zero floor
//     (doclen ∉ strides, SURVEY.md §8 a12') is applied and warp 0 adds the 32 rows.
// The box is always 16 rows high: a document's last tile also fetches up to 15 rows of its successor (masked) —
// the store's 512 zero tail rows (colbert_ranker.py:62) keep that in bounds.  (Exact-height boxes, one tensor map per
// height as in the 128-wide kernel, were tried: 6.6 vs 6.5 ms at dim 768, 8.6 vs 8.0 ms at dim 1024 — this kernel is bound
// by the per-tile barrier + reduction, not by the bytes it reads.)
#include <algorithm>

#include "cbk_common.cuh"

namespace cbk {

int make_store_tensor_map_3d_wide(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_rows);

namespace {

constexpr int kWWarps = 8;
constexpr int kWThreads = kWWarps * 32;
constexpr int kWTileRows = 16;
constexpr int kWSegCands = 32;           // candidates per claimed segment
constexpr int kWMaxKs = 8;               // k-steps (16 columns) per warp ⇒ dim ≤ 8 × 8 × 16 = 1024
constexpr int kWMaxStages = 32;          // narrow rows need many small tiles in flight
constexpr int kWPartialFloats = 16 * 32; // one warp's partial accumulator tile: 16 registers × 32 lanes

struct StrideSet {
  int n;
  int v[CBK_MAX_STRIDES];
};

__device__ __forceinline__ uint32_t bf16x2_to_f16x2_w(uint32_t v) {
  const float lo = __uint_as_float(v << 16);
  const float hi = __uint_as_float(v & 0xffff0000u);
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

struct WideCtl {                          // static shared memory
  uint64_t full[kWMaxStages];
  int2 meta[kWSegCands];                  // (first row, doclen) of the segment's candidates; doclen ≤ 0: nothing to score
  float rowmax[2][32];                    // [token half][query row] of the document just finished
  unsigned int seg;
};

// kMaxKs: k-steps per warp this instantiation can hold (dim ≤ 128·kMaxKs); kCtas: CTAs per SM it is compiled for — narrower
// rows need fewer query registers, and two resident CTAs overlap one's block barrier and reduction with the other's MMAs
template <typename T, bool kCvtBf16, int kMaxKs, int kCtas>
__global__ void __launch_bounds__(kWThreads, kCtas)
maxsim_wide_kernel(const __grid_constant__ CUtensorMap tmap, const int64_t* __restrict__ pfxsum,
                   const int32_t* __restrict__ doclens, int64_t n_docs, int64_t pid_base, int skip_foreign, StrideSet strides,
                   const float* __restrict__ Q, const int32_t* __restrict__ q_lens, int q_len, int dim, int64_t n_queries, const int64_t* __restrict__ cand_pids,
                   const int64_t* __restrict__ rowptr, int64_t n_cand_bound, int n_stages, float* __restrict__ out,
                   unsigned int* __restrict__ seg_counter) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ WideCtl ctl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  const uint32_t tile_bytes = static_cast<uint32_t>(dim) * 2u * kWTileRows;
  const uint32_t tiles_addr = raw_addr + pad;
  float* partials = reinterpret_cast<float*>(smem_raw + pad + static_cast<size_t>(n_stages) * tile_bytes);   // [2][8][512]
  const uint32_t full0 = smem_u32(&ctl.full[0]);

  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < n_stages; ++s) mbar_init(full0 + 8 * s, 1);
    fence_mbar_init();
  }

  // my slice of the k-steps
  const int n_ks_total = dim >> 4;
  const int ks_base = n_ks_total / kWWarps, ks_rem = n_ks_total % kWWarps;
  const int my_nks = ks_base + (warp < ks_rem ? 1 : 0);
  const int my_ks0 = warp * ks_base + min(warp, ks_rem);

  const int64_t n_cand = min(rowptr[n_queries], n_cand_bound);
  const int64_t n_segs = (n_cand + kWSegCands - 1) / kWSegCands;
  uint32_t p_stage = 0;                   // producer: next stage to fill
  uint32_t c_stage = 0, c_parity = 0;     // consumer: next stage to read / parity of its next full phase
  int64_t cur_q = -1;
  uint32_t qa[2][kMaxKs][4];              // A fragments of my slice of the current query: [m-tile][k-step][reg]
  int buf = 0;

  // ldmatrix.x4 of one k-step: matrices 0/1 = tokens 0-7, columns k..k+7 / k+8..k+15; matrices 2/3 = tokens 8-15
  const int lrow = ((lane >> 4) << 3) + (lane & 7);          // tile row this lane addresses
  const int lhalf = (lane >> 3) & 1;                          // low / high 8 columns of the k-step
  // reduction role: this warp adds up accumulator registers 2w', 2w'+1 = token half w'/4, query rows (w'%4)*8 + lane/4
  const int red_half = warp >> 2;
  const int red_row = ((warp & 3) << 3) + (lane >> 2);

  __syncthreads();

  while (true) {
    __syncthreads();                                           // the previous segment is finished with ctl.meta
    if (tid == 0) ctl.seg = atomicAdd(seg_counter, 1u);
    __syncthreads();
    const unsigned int seg = ctl.seg;
    if (static_cast<int64_t>(seg) >= n_segs) break;
    const int64_t c0 = static_cast<int64_t>(seg) * kWSegCands;
    const int nc = static_cast<int>(min(static_cast<int64_t>(kWSegCands), n_cand - c0));
    if (tid < nc) {
      const int64_t pid = cand_pids[c0 + tid] - pid_base;
      int2 m = make_int2(0, -1);
      if (pid < 0 || pid >= n_docs) {
        out[c0 + tid] = skip_foreign ? -INFINITY : __int_as_float(0x7fc00000);
      } else {
        m.x = static_cast<int>(pfxsum[pid]);
        m.y = doclens[pid];
        if (m.y == 0) out[c0 + tid] = 0.f;                     // an empty document scores 0 (all-false mask in the reference)
      }
      ctl.meta[tid] = m;
    }
    __syncthreads();

    // owning query of the first candidate (uniform guess, else binary search), as in the 128-wide kernel
    int64_t q = 0;
    {
      int64_t lo = 0, hi = n_queries - 1;
      while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (rowptr[mid] <= c0) lo = mid; else hi = mid - 1;
      }
      q = lo;
    }
    int64_t q_end = rowptr[q + 1];

    // producer cursor: runs n_stages - 1 tiles ahead of the consumer inside the segment (every thread keeps the same
    // books; one elected lane of warp 0 issues)
    int pc = 0, pt = 0;
    auto issue_tile = [&]() {
      while (pc < nc && ctl.meta[pc].y <= 0) ++pc;
      if (pc >= nc) return;
      const int2 m = ctl.meta[pc];
      if (warp == 0) {
        if (elect_one()) {
          const uint32_t bar = full0 + 8 * p_stage;
          mbar_arrive_expect_tx(bar, tile_bytes);
          tma_load_3d(tiles_addr + p_stage * tile_bytes, &tmap, 0, m.x + pt * kWTileRows, 0, bar, kEvictFirst);
        }
        __syncwarp();
      }
      if (++p_stage == static_cast<uint32_t>(n_stages)) p_stage = 0;
      ++pt;
      if (pt * kWTileRows >= m.y) {
        pt = 0;
        ++pc;
      }
    };
    for (int s = 0; s < n_stages - 1; ++s) issue_tile();

    for (int ci = 0; ci < nc; ++ci) {
      const int2 m = ctl.meta[ci];
      const int len = m.y;
      if (len <= 0) continue;
      const int64_t c = c0 + ci;
      while (c >= q_end) {
        ++q;
        q_end = rowptr[q + 1];
      }
      if (q != cur_q) {
        // ---- (re)load my slice of the query as A fragments, fp32 → T with round-to-nearest ----------------------
        cur_q = q;
        const float* Qq = Q + q * static_cast<int64_t>(q_len) * dim;
        const int ql = q_lens ? min(q_len, q_lens[q]) : q_len;   // rows at or past this query's own length read as zero
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r0 = mt * 16 + (lane >> 2);
          const int r1 = r0 + 8;
#pragma unroll
          for (int j = 0; j < kMaxKs; ++j) {
            if (j < my_nks) {
              const int k0 = (my_ks0 + j) * 16 + 2 * (lane & 3);
              float2 v00 = make_float2(0.f, 0.f), v10 = v00, v01 = v00, v11 = v00;
              if (r0 < ql) {
                v00 = *reinterpret_cast<const float2*>(Qq + static_cast<int64_t>(r0) * dim + k0);
                v01 = *reinterpret_cast<const float2*>(Qq + static_cast<int64_t>(r0) * dim + k0 + 8);
              }
              if (r1 < ql) {
                v10 = *reinterpret_cast<const float2*>(Qq + static_cast<int64_t>(r1) * dim + k0);
                v11 = *reinterpret_cast<const float2*>(Qq + static_cast<int64_t>(r1) * dim + k0 + 8);
              }
              qa[mt][j][0] = pack2<T>(v00.x, v00.y);
              qa[mt][j][1] = pack2<T>(v10.x, v10.y);
              qa[mt][j][2] = pack2<T>(v01.x, v01.y);
              qa[mt][j][3] = pack2<T>(v11.x, v11.y);
            }
          }
        }
      }

      const int ntiles = (len + kWTileRows - 1) / kWTileRows;
      float rmax = -INFINITY;                                   // running max of (red_half, red_row), lanes of a row agree
      for (int t = 0; t < ntiles; ++t) {
        issue_tile();
        mbar_wait(full0 + 8 * c_stage, c_parity);
        const uint32_t sbase = tiles_addr + c_stage * tile_bytes;

        float acc[2][2][4];                                     // [token half][m-tile][reg]
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[s][mt][r] = 0.f;
        // all of my B fragments first (their shared-memory latencies overlap), then the MMAs
        uint32_t bf[kMaxKs][4];
#pragma unroll
        for (int j = 0; j < kMaxKs; ++j) {
          if (j < my_nks) {
            const int ks = my_ks0 + j;
            const int chunk = ((ks & 3) << 1) + lhalf;          // 16-byte chunk inside the 128-byte slab row
            ldmatrix_x4(sbase + (ks >> 2) * (kWTileRows * 128) + lrow * 128 + (((chunk ^ lrow) & 7) << 4), bf[j][0], bf[j][1],
                        bf[j][2], bf[j][3]);
          }
        }
#pragma unroll
        for (int j = 0; j < kMaxKs; ++j) {
          if (j < my_nks) {
            uint32_t b0 = bf[j][0], b1 = bf[j][1], b2 = bf[j][2], b3 = bf[j][3];
            if (kCvtBf16) {
              b0 = bf16x2_to_f16x2_w(b0);
              b1 = bf16x2_to_f16x2_w(b1);
              b2 = bf16x2_to_f16x2_w(b2);
              b3 = bf16x2_to_f16x2_w(b3);
            }
            mma_16816<T>(acc[0][0], qa[0][j], b0, b1);
            mma_16816<T>(acc[0][1], qa[1][j], b0, b1);
            mma_16816<T>(acc[1][0], qa[0][j], b2, b3);
            mma_16816<T>(acc[1][1], qa[1][j], b2, b3);
          }
        }
        // park my partial tile: register r = half*8 + mt*4 + reg at [r*32 + lane] (conflict-free)
        float* mine = partials + (buf * kWWarps + warp) * kWPartialFloats;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int r = 0; r < 4; ++r) mine[(s * 8 + mt * 4 + r) * 32 + lane] = acc[s][mt][r];
        __syncthreads();   // all partials of this tile are parked; every warp is also done reading stage `st`

        // add up the eight partials of registers 2*warp and 2*warp+1: two adjacent tokens of one query row
        const float* pb = partials + buf * kWWarps * kWPartialFloats + (2 * warp) * 32 + lane;
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int w = 0; w < kWWarps; ++w) {
          v0 += pb[w * kWPartialFloats];
          v1 += pb[w * kWPartialFloats + 32];
        }
        const int tok = t * kWTileRows + red_half * 8 + 2 * (lane & 3);
        if (tok >= len) v0 = -INFINITY;
        if (tok + 1 >= len) v1 = -INFINITY;
        rmax = fmaxf(rmax, fmaxf(v0, v1));
        buf ^= 1;
        if (++c_stage == static_cast<uint32_t>(n_stages)) {
          c_stage = 0;
          c_parity ^= 1u;
        }
      }
      rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 1));
      rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 2));
      if ((lane & 3) == 0) ctl.rowmax[red_half][red_row] = rmax;
      __syncthreads();
      if (warp == 0) {
        bool do_floor = strides.n > 0;
#pragma unroll
        for (int i = 0; i < CBK_MAX_STRIDES; ++i)
          if (i < strides.n && strides.v[i] == len) do_floor = false;
        float v = fmaxf(ctl.rowmax[0][lane], ctl.rowmax[1][lane]);
        if (do_floor) v = fmaxf(v, 0.f);
        if (lane >= q_len) v = 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) out[c] = v;
      }
      // ctl.rowmax is rewritten only after the next document's first tile barrier: warp 0 has read it by then
    }
  }
}

template <typename T, bool kCvtBf16, int kMaxKs, int kCtas>
int launch_wide_as(const CUtensorMap& tmap, const int64_t* pfxsum, const int32_t* doclens, int64_t n_docs, int64_t pid_base,
                   int skip_foreign, const StrideSet& strides, const float* Q, const int32_t* q_lens, int q_len, int dim, int64_t n_queries,
                   const int64_t* cand_pids, const int64_t* rowptr, int64_t n_cand, float* out, unsigned int* counter,
                   cudaStream_t stream) {
  const size_t tile_bytes = static_cast<size_t>(dim) * 2 * kWTileRows;
  const size_t partial_bytes = 2 * kWWarps * kWPartialFloats * sizeof(float);
  const size_t budget = (224 * 1024) / kCtas - 2048;
  const int n_stages = static_cast<int>(std::max<size_t>(2, std::min<size_t>(kWMaxStages, (budget - partial_bytes) / tile_bytes)));
  const size_t smem = n_stages * tile_bytes + partial_bytes + 1024;
  auto kern = maxsim_wide_kernel<T, kCvtBf16, kMaxKs, kCtas>;
  CBK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t n_segs = (n_cand + kWSegCands - 1) / kWSegCands;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_segs, static_cast<int64_t>(sm_count()) * kCtas)));
  kern<<<grid, kWThreads, smem, stream>>>(tmap, pfxsum, doclens, n_docs, pid_base, skip_foreign, strides, Q, q_lens, q_len, dim, n_queries,
                                         cand_pids, rowptr, n_cand, n_stages, out, counter);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

template <typename T, bool kCvtBf16>
int launch_wide(const CUtensorMap& tmap, const int64_t* pfxsum, const int32_t* doclens, int64_t n_docs, int64_t pid_base,
                int skip_foreign, const StrideSet& strides, const float* Q, const int32_t* q_lens, int q_len, int dim, int64_t n_queries,
                const int64_t* cand_pids, const int64_t* rowptr, int64_t n_cand, float* out, unsigned int* counter,
                cudaStream_t stream) {
  if (dim <= 384)
    return launch_wide_as<T, kCvtBf16, 3, 2>(tmap, pfxsum, doclens, n_docs, pid_base, skip_foreign, strides, Q, q_lens, q_len, dim, n_queries,
                                             cand_pids, rowptr, n_cand, out, counter, stream);
  if (dim <= 768)
    return launch_wide_as<T, kCvtBf16, 6, 2>(tmap, pfxsum, doclens, n_docs, pid_base, skip_foreign, strides, Q, q_lens, q_len, dim, n_queries,
                                             cand_pids, rowptr, n_cand, out, counter, stream);
  return launch_wide_as<T, kCvtBf16, 8, 1>(tmap, pfxsum, doclens, n_docs, pid_base, skip_foreign, strides, Q, q_lens, q_len, dim, n_queries,
                                           cand_pids, rowptr, n_cand, out, counter, stream);
}

}  // namespace

bool rerank_wide_supports(int dim) { return dim % 64 == 0 && dim >= 64 && dim <= 1024; }

int rerank_wide_dispatch(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                         const int32_t* d_doclens, int64_t n_docs, int64_t pid_base, const int32_t* strides, int n_strides,
                         const float* d_Q, const int32_t* d_q_lens, int q_len, int64_t n_queries, const int64_t* d_cand_pids,
                         const int64_t* d_cand_rowptr, int64_t n_cand_total, float* d_out_scores, void* d_workspace,
                         int flags, cudaStream_t stream) {
  static thread_local CUtensorMap tmap;
  static thread_local const void* cached_base = nullptr;
  static thread_local int64_t cached_rows = -1;
  static thread_local int cached_dim = -1;
  if (cached_base != d_store || cached_rows != n_store_rows || cached_dim != dim) {
    cached_base = nullptr;
    int rc = make_store_tensor_map_3d_wide(&tmap, d_store, n_store_rows, dim, kWTileRows);
    if (rc != CBK_OK) return rc;
    cached_base = d_store;
    cached_rows = n_store_rows;
    cached_dim = dim;
  }
  StrideSet ss;
  ss.n = n_strides;
  for (int i = 0; i < CBK_MAX_STRIDES; ++i) ss.v[i] = i < n_strides ? strides[i] : -1;
  const int skip = (flags & CBK_FLAG_SKIP_FOREIGN_PIDS) ? 1 : 0;
  unsigned int* counter = static_cast<unsigned int*>(d_workspace);
  CBK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  if (store_dtype == CBK_F16)
    return launch_wide<__half, false>(tmap, d_pfxsum, d_doclens, n_docs, pid_base, skip, ss, d_Q, d_q_lens, q_len, dim, n_queries, d_cand_pids,
                                      d_cand_rowptr, n_cand_total, d_out_scores, counter, stream);
  if (flags & CBK_FLAG_BF16_NATIVE_MMA)
    return launch_wide<__nv_bfloat16, false>(tmap, d_pfxsum, d_doclens, n_docs, pid_base, skip, ss, d_Q, d_q_lens, q_len, dim, n_queries,
                                             d_cand_pids, d_cand_rowptr, n_cand_total, d_out_scores, counter, stream);
  return launch_wide<__half, true>(tmap, d_pfxsum, d_doclens, n_docs, pid_base, skip, ss, d_Q, d_q_lens, q_len, dim, n_queries, d_cand_pids,
                                   d_cand_rowptr, n_cand_total, d_out_scores, counter, stream);
}

}  // namespace cbk
