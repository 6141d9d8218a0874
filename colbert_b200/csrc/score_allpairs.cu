// All-pairs MaxSim of padded query / document batches, forward + backward (sm_100a: tcgen05 + TMEM + TMA).
//
// This is BaseModel.score as the reference TRAINS with it (colbert/modeling/BaseModel.py:39-46, called from
// colbert/modeling/colbert_model.py:87-95 on the all-gathered batch of colbert/training/training_utils.py:35-45):
//
//     scores[q, d] = Σ_m max_n (Q[q,m]·q_mask[q,m]) · (D[d,n]·d_mask[d,n])          Q [q, m, h], D [d, n, h]
//
// at the author's width h = 768 (proj_conf/dense.yaml:8), where the problem is a [q·m, h] × [h, d·n] GEMM
// (q = 170, d = 340, m = 32, n = 384: 1.09 TFLOP) whose 2.8 GB product `simmat` the reference materialises.
// Here the product never leaves the SM:
//
//   forward   A = 4 queries × 32 rows (one 128-lane TMEM accumulator), B = 256 consecutive rows of the flattened
//             documents, K streamed in 64-wide slabs through a 4-stage TMA ring (A slab 16 KB + B slab 32 KB per stage);
//             tcgen05.mma M = 128, N = 256, two 256-column accumulators in TMEM so the epilogue of one tile runs under
//             the MMAs of the next.  Epilogue: each lane (= query row) folds its 256 columns into a running
//             (max, argmax) of the open document, closes documents at their last row (every document has exactly n
//             rows), sums the 32 rows of a query with shuffles and stores score[q, d] plus argmax[q, d, m] (the
//             reference's `indices`, BaseModel.py:44, which autograd keeps for the backward).
//             Masked rows arrive as zero rows (cbk_mask_cast_rows fuses the multiplicative masks of BaseModel.py:41-42
//             with the cast to 16 bits), so a masked slot scores exactly 0 as in the reference.
//   backward  dQ[q,m] = q_mask · Σ_d g[q,d] · Dm[d, argmax[q,d,m]]          (gather-accumulate, one warp per query row)
//             dD[d,n] = d_mask · Σ_{(q,m): argmax[q,d,m] = n} g[q,d] · Qm[q,m]
//                       (one CTA per document: the (q, m) pairs are bucketed by n in shared memory — counting sort with a
//                        deterministic, ordered fill — and each warp sums the rows of its buckets in that order: no
//                        floating-point atomics, bit-reproducible gradients)
//             both read the 16-bit operands of the forward pass, i.e. they are the exact gradient of the function the
//             forward pass computed.
#include <algorithm>

#include "umma.cuh"

namespace cbk {

int make_rows_tensor_map(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_rows);
int make_query_block_tensor_map(CUtensorMap* out, const void* base, int64_t n_queries, int m, int dim, int rows_per_query);

namespace {

constexpr int kApTileN = 256;                 // document rows per accumulator
constexpr int kApStages = 4;
constexpr int kApABytes = 128 * 128;          // one K slab (64 × 16 bit) of the query block
constexpr int kApBBytes = kApTileN * 128;     // one K slab of the document tile
constexpr int kApStageBytes = kApABytes + kApBBytes;
constexpr int kApThreads = 6 * 32;            // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue

struct ApMaps {
  CUtensorMap q;   // packed queries as {h, m, q}: box {64, rpq, 128 / rpq} = one 128-row block of 4 queries x 32 row slots (or, for
                   // m <= 16, 8 queries x 16), rows >= m / queries >= q zero-filled
  CUtensorMap d;   // packed documents as [d*n, h]: box {64, 256}, rows past the end zero-filled
};

// One accumulator column: x = similarity of this lane's query row with document row `row` (row index inside the item).
#define CBK_AP_FOLD(x, row)          \
  {                                  \
    const float _x = (x);            \
    if (_x > best) {                 \
      best = _x;                     \
      bi = (row);                    \
    }                                \
  }

__global__ void __launch_bounds__(kApThreads, 1)
score_allpairs_fwd_kernel(const __grid_constant__ ApMaps maps, int nq, int m, int rpq, int nd, int n, int n_slabs, int dpi,
                          int n_qblocks, int n_items, uint32_t idesc, float* __restrict__ scores,
                          int32_t* __restrict__ argmax) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kApStages], bar_empty[kApStages];
  __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (tid == 0) {
    for (int s = 0; s < kApStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), 4);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(smem_u32(&tmem_base_smem), 512);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;

  // item i = (document range r = i / n_qblocks, query block qb = i % n_qblocks): query blocks vary fastest, so the CTAs
  // running at any moment share a handful of document ranges (L2) and sweep all query blocks over them.
  auto item_docs = [&](int i, int& d0, int& nd_i) {
    const int r = i / n_qblocks;
    d0 = r * dpi;
    nd_i = min(dpi, nd - d0);
  };

  if (warp == 0) {
    // ===================================== TMA producer =============================================
    if (lane == 0) {
      tma_prefetch_desc(&maps.q);
      tma_prefetch_desc(&maps.d);
    }
    uint32_t it = 0;
    for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
      int d0, nd_i;
      item_docs(i, d0, nd_i);
      const int qb = i % n_qblocks;
      const int64_t row0 = static_cast<int64_t>(d0) * n;
      const int n_tiles = (nd_i * n + kApTileN - 1) / kApTileN;
      for (int t = 0; t < n_tiles; ++t)
        for (int ks = 0; ks < n_slabs; ++ks) {
          const uint32_t st = it % kApStages;
          mbar_wait(smem_u32(&bar_empty[st]), ((it / kApStages) & 1u) ^ 1u);
          const uint32_t full = smem_u32(&bar_full[st]);
          const uint32_t dst = stage0 + st * kApStageBytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(full, kApStageBytes);
            tma_load_3d(dst, &maps.q, ks * 64, 0, qb * (128 / rpq), full, kEvictLast);
            tma_load_2d(dst + kApABytes, &maps.d, ks * 64, static_cast<int>(row0 + static_cast<int64_t>(t) * kApTileN), full,
                        kEvictNormal);
          }
          __syncwarp();
          ++it;
        }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ===============================================
    const uint32_t full0 = hold(smem_u32(&bar_full[0])), empty0 = hold(smem_u32(&bar_empty[0]));
    const uint32_t accfull0 = hold(smem_u32(&bar_acc_full[0])), accempty0 = hold(smem_u32(&bar_acc_empty[0]));
    const uint32_t a_lo0 = hold(umma::desc_lo_sw128(stage0)), b_lo0 = hold(umma::desc_lo_sw128(stage0 + kApABytes));
    constexpr uint32_t kStageDesc = kApStageBytes >> 4;
    uint32_t st = 0, st_parity = 0, acc_it = 0;
    for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
      int d0, nd_i;
      item_docs(i, d0, nd_i);
      const int n_tiles = (nd_i * n + kApTileN - 1) / kApTileN;
      for (int t = 0; t < n_tiles; ++t, ++acc_it) {
        const uint32_t slot = acc_it & 1u;
        mbar_wait(accempty0 + 8 * slot, ((acc_it >> 1) & 1u) ^ 1u);
        umma::fence_after_sync();
        const uint32_t d_tmem = tmem + slot * kApTileN;
        for (int ks = 0; ks < n_slabs; ++ks) {
          mbar_wait(full0 + 8 * st, st_parity);
          umma::fence_after_sync();
          const uint32_t a_lo = a_lo0 + st * kStageDesc, b_lo = b_lo0 + st * kStageDesc;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma::mma_f16_ss_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (ks | k) ? 1u : 0u);
            umma::commit(empty0 + 8 * st);
          }
          __syncwarp();
          if (++st == kApStages) {
            st = 0;
            st_parity ^= 1u;
          }
        }
        if (elect_one()) umma::commit(accfull0 + 8 * slot);
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue (warps 2..5) ====================================
    const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read: 32 accumulator rows =
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;   // one query (rpq = 32) or two (rpq = 16: multi-view shapes)
    const int row_in_q = lane & (rpq - 1);                       // this lane's query row
    uint32_t acc_it = 0;
    for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
      int d0, nd_i;
      item_docs(i, d0, nd_i);
      const int qb = i % n_qblocks;
      const int q = qb * (128 / rpq) + quad * (32 / rpq) + lane / rpq;     // this lane's query
      const bool write = q < nq;
      const int n_rows = nd_i * n;
      const int n_tiles = (n_rows + kApTileN - 1) / kApTileN;
      float best = -INFINITY;
      int bi = 0;
      int doc = d0;            // open document
      int next_end = n - 1;    // its last row (row index inside the item)
      for (int t = 0; t < n_tiles; ++t, ++acc_it) {
        const uint32_t slot = acc_it & 1u;
        mbar_wait(smem_u32(&bar_acc_full[slot]), (acc_it >> 1) & 1u);
        umma::fence_after_sync();
        const uint32_t t_addr = tmem + lane_base + slot * kApTileN;
#pragma unroll 1
        for (int c = 0; c < kApTileN / 32; ++c) {
          uint32_t v[32];
          umma::tmem_ld_32x32(t_addr + c * 32, v);
          umma::tmem_ld_wait();
          if (c == kApTileN / 32 - 1) {   // the whole accumulator is in registers or folded: hand the slot back
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[slot]));
          }
          const int r0 = t * kApTileN + c * 32;
          if (r0 >= n_rows) continue;     // rows of the next item (or past the end of D): never part of an open document
          if (next_end >= r0 + 32) {
            // no document ends inside these 32 columns: four independent (max, argmax) chains, merged keeping the
            // FIRST maximal column (torch.max's tie rule on the CPU reference)
            float b4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int i4[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = __uint_as_float(v[j]);
              if (x > b4[j & 3]) {
                b4[j & 3] = x;
                i4[j & 3] = j;
              }
            }
#pragma unroll
            for (int k = 1; k < 4; ++k)
              if (b4[k] > b4[0] || (b4[k] == b4[0] && i4[k] < i4[0])) {
                b4[0] = b4[k];
                i4[0] = i4[k];
              }
            CBK_AP_FOLD(b4[0], r0 + i4[0]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              CBK_AP_FOLD(__uint_as_float(v[j]), r0 + j);
              if (r0 + j == next_end) {   // warp-uniform: last row of the open document
                float s = best;
                if (rpq == 32) s += __shfl_xor_sync(0xffffffffu, s, 16);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (write) {
                  const int64_t o = static_cast<int64_t>(q) * nd + doc;
                  if (row_in_q == 0) scores[o] = s;
                  if (argmax != nullptr && row_in_q < m) argmax[o * m + row_in_q] = bi - (next_end - (n - 1));
                }
                ++doc;
                best = -INFINITY;
                bi = next_end + 1;
                next_end += n;
                if (next_end >= n_rows) next_end = 0x7fffffff;   // past the item's last document: the rest of the tile belongs to nobody
              }
            }
          }
        }
      }
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// ---- backward ---------------------------------------------------------------------------------------

template <typename T>
__device__ __forceinline__ void fma8(float (&acc)[8], const uint4& u, float g);

template <>
__device__ __forceinline__ void fma8<__half>(float (&acc)[8], const uint4& u, float g) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    acc[2 * j] = fmaf(g, f.x, acc[2 * j]);
    acc[2 * j + 1] = fmaf(g, f.y, acc[2 * j + 1]);
  }
}
template <>
__device__ __forceinline__ void fma8<__nv_bfloat16>(float (&acc)[8], const uint4& u, float g) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    acc[2 * j] = fmaf(g, f.x, acc[2 * j]);
    acc[2 * j + 1] = fmaf(g, f.y, acc[2 * j + 1]);
  }
}

template <typename M>
__device__ __forceinline__ float mask_value(const void* mask, int64_t i) {
  return mask ? static_cast<float>(static_cast<const M*>(mask)[i]) : 1.f;
}
__device__ __forceinline__ float mask_at(const void* mask, int mask_dtype, int64_t i) {
  switch (mask_dtype) {
    case CBK_MASK_U8: return mask_value<uint8_t>(mask, i);
    case CBK_MASK_I64: return mask_value<int64_t>(mask, i);
    case CBK_MASK_F32: return mask_value<float>(mask, i);
    default: return 1.f;
  }
}

// A lane owns 8 consecutive columns of every 256-column chunk of a row: chunk c covers columns 256 c + 8 lane .. + 7.
template <int HC>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&acc)[HC][8], float scale, int dim, int lane) {
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < dim) {
      float4* o = reinterpret_cast<float4*>(dst + col);
      o[0] = make_float4(acc[c][0] * scale, acc[c][1] * scale, acc[c][2] * scale, acc[c][3] * scale);
      o[1] = make_float4(acc[c][4] * scale, acc[c][5] * scale, acc[c][6] * scale, acc[c][7] * scale);
    }
  }
}

constexpr int kDqRows = 8;      // query rows (warps) per CTA
constexpr int kDqDocs = 32;     // documents staged per step

// dQ[q, mm, :] = q_mask[q, mm] * Σ_d g[q, d] * Dp[d*n + argmax[q, d, mm], :]
template <typename T, int HC>
__global__ void __launch_bounds__(kDqRows * 32)
score_allpairs_bwd_dq_kernel(const T* __restrict__ Dp, const int32_t* __restrict__ argmax, const float* __restrict__ G,
                             const void* __restrict__ q_mask, int q_mask_dtype, int nq, int m, int nd, int n, int dim,
                             float* __restrict__ dQ) {
  __shared__ int idx_s[kDqDocs][kDqRows];
  __shared__ float g_s[kDqDocs];
  const int m_groups = (m + kDqRows - 1) / kDqRows;
  const int q = blockIdx.x / m_groups, m0 = (blockIdx.x % m_groups) * kDqRows;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mm = m0 + warp;
  float acc[HC][8];
#pragma unroll
  for (int c = 0; c < HC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;

  for (int d0 = 0; d0 < nd; d0 += kDqDocs) {
    const int cnt = min(kDqDocs, nd - d0);
    {
      const int dd = tid / kDqRows, j = tid % kDqRows;
      int v = 0;
      if (dd < cnt && m0 + j < m) v = argmax[(static_cast<int64_t>(q) * nd + d0 + dd) * m + m0 + j];
      idx_s[dd][j] = v;
      if (tid < kDqDocs) g_s[tid] = tid < cnt ? G[static_cast<int64_t>(q) * nd + d0 + tid] : 0.f;
    }
    __syncthreads();
    if (mm < m) {
#pragma unroll 4
      for (int dd = 0; dd < cnt; ++dd) {
        const float g = g_s[dd];
        const uint4* row = reinterpret_cast<const uint4*>(Dp + (static_cast<int64_t>(d0 + dd) * n + idx_s[dd][warp]) * dim);
#pragma unroll
        for (int c = 0; c < HC; ++c)
          if (c * 256 + lane * 8 < dim) fma8<T>(acc[c], __ldg(row + c * 32 + lane), g);
      }
    }
    __syncthreads();
  }
  if (mm < m) {
    const int64_t r = static_cast<int64_t>(q) * m + mm;
    store_row<HC>(dQ + r * dim, acc, mask_at(q_mask, q_mask_dtype, r), dim, lane);
  }
}

constexpr int kDdThreads = 512;

// dD[d, b, :] = d_mask[d, b] * Σ_{e = (q, mm): argmax[q, d, mm] = b} g[q, d] * Qp[e, :],   entries summed in ascending e
template <typename T, int HC>
__global__ void __launch_bounds__(kDdThreads)
score_allpairs_bwd_dd_kernel(const T* __restrict__ Qp, const int32_t* __restrict__ argmax, const float* __restrict__ G,
                             const void* __restrict__ d_mask, int d_mask_dtype, int nq, int m, int nd, int n, int dim,
                             float* __restrict__ dD) {
  extern __shared__ uint8_t smem_raw[];
  const int E = nq * m;                                   // (q, mm) pairs, entry e = q*m + mm = row of Qp
  const int E8 = (E + 7) & ~7;
  uint16_t* idx_s = reinterpret_cast<uint16_t*>(smem_raw);             // [E8] bucket of every entry
  uint16_t* list_s = idx_s + E8;                                        // [E8] entries grouped by bucket, ascending inside a bucket
  int* cnt_s = reinterpret_cast<int*>(list_s + E8);                     // [n + 1]
  int* off_s = cnt_s + (n + 1);                                         // [n + 1]
  __shared__ int next_bucket;
  const int d = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int b = tid; b <= n; b += kDdThreads) cnt_s[b] = 0;
  if (tid == 0) next_bucket = 0;
  __syncthreads();
  for (int e = tid; e < E8; e += kDdThreads) {
    int a = 0xffff;                                       // padding entries match no bucket (n <= 65535 → bucket ids < 0xffff)
    if (e < E) {
      const int q = e / m, mm = e - q * m;
      a = argmax[(static_cast<int64_t>(q) * nd + d) * m + mm];
      atomicAdd(&cnt_s[a], 1);
    }
    idx_s[e] = static_cast<uint16_t>(a);
  }
  __syncthreads();
  if (warp == 0) {   // exclusive scan of the bucket sizes
    int carry = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
      const int b = b0 + lane;
      const int c = b < n ? cnt_s[b] : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (b < n) off_s[b] = carry + incl - c;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) off_s[n] = carry;
  }
  __syncthreads();
  // ordered fill.  Many buckets (n > 64): the thread that owns bucket b walks every entry in ascending order and appends its
  // own.  Few buckets (multi-view shapes, n = d_view): one warp per bucket walks the entries 32 at a time and compacts the
  // matches with a ballot — still ascending, and every thread of the CTA is busy instead of n of them.
  if (n > 64) {
    for (int b = tid; b < n; b += kDdThreads) {
      int pos = off_s[b];
      if (cnt_s[b] == 0) continue;
      const uint4* p = reinterpret_cast<const uint4*>(idx_s);
      for (int e8 = 0; e8 < E8 / 8; ++e8) {
        const uint4 u = p[e8];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (static_cast<int>(w[k] & 0xffffu) == b) list_s[pos++] = static_cast<uint16_t>(e8 * 8 + 2 * k);
          if (static_cast<int>(w[k] >> 16) == b) list_s[pos++] = static_cast<uint16_t>(e8 * 8 + 2 * k + 1);
        }
      }
    }
  } else {
    for (int b = warp; b < n; b += kDdThreads / 32) {
      int pos = off_s[b];
      if (cnt_s[b] == 0) continue;
      for (int e0 = 0; e0 < E8; e0 += 32) {
        const int e = e0 + lane;
        const bool hit = e < E8 && static_cast<int>(idx_s[e]) == b;
        const uint32_t mask = __ballot_sync(0xffffffffu, hit);
        if (hit) list_s[pos + __popc(mask & ((1u << lane) - 1u))] = static_cast<uint16_t>(e);
        pos += __popc(mask);
      }
    }
  }
  __syncthreads();
  // every warp takes buckets from a shared counter (bucket sizes are very uneven) and sums its rows in list order
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(&next_bucket, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= n) break;
    float acc[HC][8];
#pragma unroll
    for (int c = 0; c < HC; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
    const int beg = off_s[b], end = beg + cnt_s[b];
#pragma unroll 4
    for (int k = beg; k < end; ++k) {
      const int e = list_s[k];
      const float g = __ldg(G + static_cast<int64_t>(e / m) * nd + d);
      const uint4* row = reinterpret_cast<const uint4*>(Qp + static_cast<int64_t>(e) * dim);
#pragma unroll
      for (int c = 0; c < HC; ++c)
        if (c * 256 + lane * 8 < dim) fma8<T>(acc[c], __ldg(row + c * 32 + lane), g);
    }
    const int64_t r = static_cast<int64_t>(d) * n + b;
    store_row<HC>(dD + r * dim, acc, mask_at(d_mask, d_mask_dtype, r), dim, lane);
  }
}

size_t dd_smem_bytes(int64_t nq, int m, int n) {
  const size_t E8 = (static_cast<size_t>(nq) * m + 7) & ~static_cast<size_t>(7);
  return E8 * 2 * sizeof(uint16_t) + 2 * static_cast<size_t>(n + 1) * sizeof(int);
}

// documents per work item: the smallest count whose rows fill whole 256-row tiles to within 3 % (or ≥ 2048 rows)
int docs_per_item(int nd, int n) {
  int best_k = 1;
  double best_waste = 1e9;
  for (int k = 1; k <= nd; ++k) {
    const int64_t rows = static_cast<int64_t>(k) * n;
    const int64_t tiles = (rows + kApTileN - 1) / kApTileN;
    const double waste = static_cast<double>(tiles * kApTileN - rows) / static_cast<double>(rows);
    if (waste < best_waste - 1e-12) {
      best_waste = waste;
      best_k = k;
    }
    if (waste <= 0.03 || rows >= 4096) break;
  }
  return best_k;
}

}  // namespace

int score_allpairs_fwd_dispatch(const void* d_Qp, const void* d_Dp, int dtype, int64_t nq, int m, int64_t nd, int n, int dim,
                                float* d_scores, int32_t* d_argmax, cudaStream_t stream) {
  ApMaps maps;
  const int rpq = m <= 16 ? 16 : 32;                  // accumulator rows per query: short (multi-view) queries share a quadrant
  int rc = make_query_block_tensor_map(&maps.q, d_Qp, nq, m, dim, rpq);
  if (rc != CBK_OK) return rc;
  rc = make_rows_tensor_map(&maps.d, d_Dp, nd * n, dim, kApTileN);
  if (rc != CBK_OK) return rc;
  const int qpb = 128 / rpq;
  const int n_qblocks = static_cast<int>((nq + qpb - 1) / qpb);
  const int dpi = docs_per_item(static_cast<int>(nd), n);
  const int n_ranges = static_cast<int>((nd + dpi - 1) / dpi);
  const int64_t n_items64 = static_cast<int64_t>(n_ranges) * n_qblocks;
  CBK_CHECK_SUPPORTED(n_items64 < (1ll << 31), "cbk_score_allpairs_fwd: too many (query block, document range) items");
  const int n_items = static_cast<int>(n_items64);
  const uint32_t fmt = dtype == CBK_BF16 ? umma::kFmtBF16 : umma::kFmtF16;
  const uint32_t idesc = umma::make_idesc(128, kApTileN, fmt, fmt);
  const size_t smem = 1024 + static_cast<size_t>(kApStages) * kApStageBytes;
  CBK_CUDA(cudaFuncSetAttribute(score_allpairs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = std::min(n_items, sm_count());
  score_allpairs_fwd_kernel<<<grid, kApThreads, smem, stream>>>(maps, static_cast<int>(nq), m, rpq, static_cast<int>(nd), n, dim / 64, dpi,
                                                              n_qblocks, n_items, idesc, d_scores, d_argmax);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

template <typename T, int HC>
static int bwd_launch(const void* d_Qp, const void* d_Dp, int64_t nq, int m, int64_t nd, int n, int dim, const float* d_G,
                      const int32_t* d_argmax, const void* d_q_mask, int q_mask_dtype, const void* d_d_mask, int d_mask_dtype,
                      float* d_dQ, float* d_dD, cudaStream_t stream) {
  if (d_dQ) {
    const int m_groups = (m + kDqRows - 1) / kDqRows;
    score_allpairs_bwd_dq_kernel<T, HC><<<static_cast<unsigned int>(nq * m_groups), kDqRows * 32, 0, stream>>>(
        static_cast<const T*>(d_Dp), d_argmax, d_G, d_q_mask, q_mask_dtype, static_cast<int>(nq), m, static_cast<int>(nd), n, dim, d_dQ);
    CBK_CUDA(cudaGetLastError());
    count_launch();
  }
  if (d_dD) {
    const size_t smem = dd_smem_bytes(nq, m, n);
    CBK_CUDA(cudaFuncSetAttribute(score_allpairs_bwd_dd_kernel<T, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    score_allpairs_bwd_dd_kernel<T, HC><<<static_cast<unsigned int>(nd), kDdThreads, smem, stream>>>(
        static_cast<const T*>(d_Qp), d_argmax, d_G, d_d_mask, d_mask_dtype, static_cast<int>(nq), m, static_cast<int>(nd), n, dim, d_dD);
    CBK_CUDA(cudaGetLastError());
    count_launch();
  }
  return CBK_OK;
}

bool score_allpairs_bwd_fits(int64_t nq, int m, int n) { return dd_smem_bytes(nq, m, n) <= 200 * 1024 && nq * m <= 65535 && n < 65535; }

int score_allpairs_bwd_dispatch(const void* d_Qp, const void* d_Dp, int dtype, int64_t nq, int m, int64_t nd, int n, int dim,
                                const float* d_G, const int32_t* d_argmax, const void* d_q_mask, int q_mask_dtype,
                                const void* d_d_mask, int d_mask_dtype, float* d_dQ, float* d_dD, cudaStream_t stream) {
  const int hc = (dim + 255) / 256;
#define CBK_BWD(HC)                                                                                                              \
  if (hc == HC) {                                                                                                                \
    if (dtype == CBK_BF16)                                                                                                       \
      return bwd_launch<__nv_bfloat16, HC>(d_Qp, d_Dp, nq, m, nd, n, dim, d_G, d_argmax, d_q_mask, q_mask_dtype, d_d_mask,        \
                                           d_mask_dtype, d_dQ, d_dD, stream);                                                    \
    return bwd_launch<__half, HC>(d_Qp, d_Dp, nq, m, nd, n, dim, d_G, d_argmax, d_q_mask, q_mask_dtype, d_d_mask, d_mask_dtype, \
                                  d_dQ, d_dD, stream);                                                                           \
  }
  CBK_BWD(1)
  CBK_BWD(2)
  CBK_BWD(3)
  CBK_BWD(4)
#undef CBK_BWD
  set_error("cbk_score_allpairs_bwd: dim %d outside (0, 1024]", dim);
  return CBK_ERR_UNSUPPORTED;
}

}  // namespace cbk
