// Multi-view MaxSim rerank at the author's own operating point (sm_100a: tcgen05 + TMEM + TMA): every document is
// exactly 16 view embeddings of the un-projected width (reference proj_conf/dense.yaml:8,29-32 — enable_multiview,
// q_view = d_view = 16, dim 768; BaseModel.get_representation BaseModel.py:21-27), every query at most 16.  Same contract
// as cbk_maxsim_rerank (reference colbert_ranker.py:88-126 + BaseModel.py:39-46) for CBK_FLAG_FIXED_DOCLEN stores.
//
// A candidate is 16 rows × dim × 2 B (24 KB at 768) read once: 16 FLOP per byte, HBM-bound.  The K-split mma.sync kernel
// (rerank_wide.cu) pays a block barrier and an 8-way shared-memory reduction per 16-row tile and reaches 0.58 of the copy
// peak here.  This kernel streams instead:
//   * B operand = 8 candidates of ONE query × 16 rows = 128 rows, one 64-column K slab per ring stage (16 KB: eight
//     2-KB TMA boxes, row = (pid − pid_base)·16, neither pfxsum nor doclens is read);
//   * A operand = the query, resident in shared memory for as long as the CTA stays on it: dim/64 slabs × 64 rows × 128 B,
//     written by a dedicated warp straight from the fp32 query in the tensor core's K-major 128-byte-swizzled layout
//     (rows 0–15 = the query rounded to the store's type; for bf16 stores rows 16–31 = the bf16 RESIDUAL of that
//     rounding, so the product keeps 16 significant bits of the query at no extra MMA; rows 32–63 zero);
//   * tcgen05.mma M = 64, N = 128, K = 16 × dim/16 into one of two 128-column TMEM accumulators; with M = 64 accumulator
//     row m lives in TMEM lane (m mod 16) + 32·(m div 16): the query rows sit in lanes 0–15 of quadrant 0, the residual
//     rows in lanes 0–15 of quadrant 1;
//   * epilogue = the two warps that own those quadrants: warp 1 passes its partial products through shared memory, warp 0
//     adds them, takes the max over each candidate's 16 columns and the sum over the 16 query rows, and writes the score
//     at the candidate's own position.
// Warp roles: 0, 1 epilogue · 2 MMA issuer · 3 query loader · 4-7 TMA producers (producer p fetches candidates p and p + 4
// of every stage: with a single producer warp the box issue rate, not HBM, set the pace — 0.75 of the copy peak).  Candidates are split into equal contiguous
// ranges over the CTAs (a CTA changes query once per ~1000 candidates; the issuer releases the A region with a
// tcgen05.commit and the loader refills it while the producer keeps the ring full).
#include <algorithm>

#include "umma.cuh"

namespace cbk {

int make_store_tensor_map(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_cols, int box_rows);

namespace {

constexpr int kMvRows = 16;                    // rows per document (d_view) and query rows at most (q_view)
constexpr int kMvTileDocs = 8;                 // candidates per tile: 8 × 16 = 128 accumulator columns
constexpr int kMvBoxBytes = kMvRows * 128;     // one candidate, one K slab
constexpr int kMvStageBytes = kMvTileDocs * kMvBoxBytes;   // 16 KB
constexpr int kMvASlabBytes = 64 * 128;        // 8 KB: 64 rows of one K slab
constexpr int kMvProducers = 4;                // TMA-issuing warps: one warp issues a 2-KB box every ~100 cycles, a tile needs 96
constexpr int kMvThreads = (4 + kMvProducers) * 32;
constexpr int kMvLoBufBytes = 128 * 16 * 4;    // residual products of one tile: [column][query row] fp32

struct MvTile {          // walk of this CTA's candidate range in tiles of up to 8 candidates of one query
  int64_t c, c_hi, q, q_end;
  const int64_t* rowptr;
  __device__ __forceinline__ void init(const int64_t* rp, int64_t n_queries, int64_t lo, int64_t hi) {
    rowptr = rp;
    c = lo;
    c_hi = hi;
    int64_t a = 0, b = n_queries - 1;      // last query whose list starts at or before lo
    while (a < b) {
      const int64_t mid = (a + b + 1) >> 1;
      if (rp[mid] <= lo) a = mid; else b = mid - 1;
    }
    q = a;
    q_end = rp[q + 1];
  }
  // → number of candidates in the next tile (0: done); c / q describe it until advance()
  __device__ __forceinline__ int next() {
    if (c >= c_hi) return 0;
    while (c >= q_end) {
      ++q;
      q_end = rowptr[q + 1];
    }
    return static_cast<int>(min(static_cast<int64_t>(kMvTileDocs), min(q_end, c_hi) - c));
  }
  __device__ __forceinline__ void advance(int n) { c += n; }
};

template <typename T>
__device__ __forceinline__ void split_hi_lo(float x, T& hi, T& lo);
template <>
__device__ __forceinline__ void split_hi_lo<__half>(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(0.f);          // fp16 stores: the query is rounded to fp16 as in every other fp16 path
}
template <>
__device__ __forceinline__ void split_hi_lo<__nv_bfloat16>(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

template <typename T>
__global__ void __launch_bounds__(kMvThreads, 1)
maxsim_mv_wide_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_docs, int64_t pid_base, int skip_foreign,
                      const float* __restrict__ Q, const int32_t* __restrict__ q_lens, int q_len, int dim, int64_t n_queries,
                      const int64_t* __restrict__ cand_pids, const int64_t* __restrict__ rowptr, int64_t n_cand_bound,
                      int n_stages, uint32_t idesc, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[16], bar_empty[16];
  __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2], bar_a_full, bar_a_free;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_slabs = dim >> 6;
  const uint32_t a_addr = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_addr = a_addr + static_cast<uint32_t>(n_slabs) * kMvASlabBytes;
  uint8_t* const smem_al = smem_raw + (a_addr - smem_u32(smem_raw));
  float* const lo_buf = reinterpret_cast<float*>(smem_al + static_cast<size_t>(n_slabs) * kMvASlabBytes +
                                                 static_cast<size_t>(n_stages) * kMvStageBytes);

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), 2);   // both epilogue warps
    }
    mbar_init(smem_u32(&bar_a_full), 1);
    mbar_init(smem_u32(&bar_a_free), 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    umma::tmem_alloc(smem_u32(&tmem_base_smem), 256);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;

  // this CTA's contiguous candidate range (a multiple of 8 candidates per CTA)
  const int64_t n_cand = min(rowptr[n_queries], n_cand_bound);
  int64_t per = (n_cand + gridDim.x - 1) / gridDim.x;
  per = (per + kMvTileDocs - 1) / kMvTileDocs * kMvTileDocs;
  const int64_t c_lo = min(n_cand, per * blockIdx.x), c_hi = min(n_cand, per * (blockIdx.x + 1));
  MvTile tl;
  tl.init(rowptr, n_queries, c_lo, c_hi);

  if (warp >= 4) {
    // ===================================== TMA producers ============================================
    const int p = warp - 4;
    if (tid == 4 * 32) tma_prefetch_desc(&tmap);
    uint32_t it = 0;
    for (int n = tl.next(); n > 0; tl.advance(n), n = tl.next()) {
      // first store row of my candidates of the tile: lane l holds candidate p + 4 l (a pid outside this store reads
      // row 0; its score is overwritten by the epilogue)
      int row = 0;
      const int my_j = p + kMvProducers * lane;
      if (my_j < n) {
        const int64_t pid = cand_pids[tl.c + my_j] - pid_base;
        row = (pid >= 0 && pid < n_docs) ? static_cast<int>(pid) * kMvRows : 0;
      }
      const int n_mine = n > p ? (n - p + kMvProducers - 1) / kMvProducers : 0;
      for (int s = 0; s < n_slabs; ++s, ++it) {
        const uint32_t st = it % static_cast<uint32_t>(n_stages);
        mbar_wait(smem_u32(&bar_empty[st]), ((it / static_cast<uint32_t>(n_stages)) & 1u) ^ 1u);
        const uint32_t full = smem_u32(&bar_full[st]);
        const uint32_t dst = b_addr + st * kMvStageBytes;
        // producer 0 arms the barrier with the bytes of the WHOLE stage; the other producers' boxes may land before
        // that (the transaction count goes negative meanwhile, the phase cannot complete before the arrival)
        if (p == 0) {
          if (elect_one()) mbar_arrive_expect_tx(full, static_cast<uint32_t>(n) * kMvBoxBytes);
          __syncwarp();
        }
        // one 2-KB box per candidate (issued by one elected lane from a converged warp: a bare UTMALDG each)
        for (int l = 0; l < n_mine; ++l) {
          const int rj = __shfl_sync(0xffffffffu, row, l);
          if (elect_one()) tma_load_2d(dst + (p + kMvProducers * l) * kMvBoxBytes, &tmap, s * 64, rj, full, kEvictFirst);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===================================== MMA issuer ===============================================
    const uint32_t full0 = hold(smem_u32(&bar_full[0])), empty0 = hold(smem_u32(&bar_empty[0]));
    const uint32_t a_lo0 = hold(umma::desc_lo_sw128(a_addr)), b_lo0 = hold(umma::desc_lo_sw128(b_addr));
    constexpr uint32_t kStageDesc = kMvStageBytes >> 4, kSlabDesc = kMvASlabBytes >> 4;
    uint32_t st = 0, st_parity = 0, acc_it = 0, n_q_seen = 0;
    int64_t cur_q = -1;
    for (int n = tl.next(); n > 0; tl.advance(n), n = tl.next(), ++acc_it) {
      if (tl.q != cur_q) {
        // every MMA issued so far read the old query: its completion frees the A region for the loader
        if (cur_q >= 0) {
          if (elect_one()) umma::commit(smem_u32(&bar_a_free));
          __syncwarp();
        }
        cur_q = tl.q;
        mbar_wait(smem_u32(&bar_a_full), n_q_seen & 1u);
        ++n_q_seen;
        umma::fence_after_sync();
      }
      const uint32_t slot = acc_it & 1u;
      mbar_wait(smem_u32(&bar_acc_empty[slot]), ((acc_it >> 1) & 1u) ^ 1u);
      umma::fence_after_sync();
      const uint32_t d_tmem = tmem + slot * 128;
      for (int s = 0; s < n_slabs; ++s) {
        mbar_wait(full0 + 8 * st, st_parity);
        umma::fence_after_sync();
        const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(s) * kSlabDesc, b_lo = b_lo0 + st * kStageDesc;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma::mma_f16_ss_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, (s | k) ? 1u : 0u);
          umma::commit(empty0 + 8 * st);
        }
        __syncwarp();
        if (++st == static_cast<uint32_t>(n_stages)) {
          st = 0;
          st_parity ^= 1u;
        }
      }
      if (elect_one()) umma::commit(smem_u32(&bar_acc_full[slot]));
      __syncwarp();
    }
  } else if (warp == 3) {
    // ===================================== query loader =============================================
    // rows 32-63 of every slab stay zero for the whole launch
    for (int i = lane; i < n_slabs * 32 * 8; i += 32) {
      const int s = i / (32 * 8), r = 32 + (i / 8) % 32, c16 = i % 8;
      *reinterpret_cast<uint4*>(smem_al + s * kMvASlabBytes + r * 128 + c16 * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    uint32_t n_q_seen = 0;
    int64_t cur_q = -1;
    for (int n = tl.next(); n > 0; tl.advance(n), n = tl.next()) {
      if (tl.q == cur_q) continue;
      cur_q = tl.q;
      if (n_q_seen > 0) mbar_wait(smem_u32(&bar_a_free), (n_q_seen - 1) & 1u);    // the MMAs of the previous query are done
      const float* Qq = Q + cur_q * static_cast<int64_t>(q_len) * dim;
      const int ql = q_lens ? min(q_len, q_lens[cur_q]) : q_len;
      // one 16-byte chunk (8 columns) per step: (row r, slab s, chunk c16) → row r (rounded value) and row 16 + r (residual)
      const int n_chunks = kMvRows * n_slabs * 8;
      for (int i = lane; i < n_chunks; i += 32) {
        const int r = i / (n_slabs * 8), s = (i / 8) % n_slabs, c16 = i % 8;
        uint4 hi4 = make_uint4(0u, 0u, 0u, 0u), lo4 = hi4;
        if (r < ql) {
          const float4 f0 = *reinterpret_cast<const float4*>(Qq + static_cast<int64_t>(r) * dim + s * 64 + c16 * 8);
          const float4 f1 = *reinterpret_cast<const float4*>(Qq + static_cast<int64_t>(r) * dim + s * 64 + c16 * 8 + 4);
          const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          T h[8], l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_hi_lo<T>(x[j], h[j], l[j]);
          hi4 = *reinterpret_cast<const uint4*>(h);
          lo4 = *reinterpret_cast<const uint4*>(l);
        }
        // K-major, 128-byte swizzle: the 16-byte chunk index is XORed with the row's low three bits
        uint8_t* slab = smem_al + s * kMvASlabBytes;
        *reinterpret_cast<uint4*>(slab + r * 128 + ((c16 ^ (r & 7)) << 4)) = hi4;
        *reinterpret_cast<uint4*>(slab + (16 + r) * 128 + ((c16 ^ (r & 7)) << 4)) = lo4;
      }
      fence_proxy_async();            // generic-proxy writes → visible to the tensor core's (async-proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a_full));
      ++n_q_seen;
    }
  } else {
    // ===================================== epilogue (warps 0 and 1) =================================
    // warp 0 = TMEM quadrant 0 (query rows × columns), warp 1 = quadrant 1 (residual rows); lanes 16-31 hold nothing
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t acc_it = 0;
    for (int n = tl.next(); n > 0; tl.advance(n), n = tl.next(), ++acc_it) {
      const uint32_t slot = acc_it & 1u;
      mbar_wait(smem_u32(&bar_acc_full[slot]), (acc_it >> 1) & 1u);
      umma::fence_after_sync();
      const uint32_t t_addr = tmem + lane_base + slot * 128;
      if (warp == 1) {
#pragma unroll 1
        for (int j = 0; j < kMvTileDocs; ++j) {
          uint32_t v[16];
          umma::tmem_ld_32x16(t_addr + j * 16, v);
          umma::tmem_ld_wait();
          if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) lo_buf[(j * 16 + i) * 16 + lane] = __uint_as_float(v[i]);
          }
        }
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[slot]));
        asm volatile("bar.sync 1, 64;" ::: "memory");     // residual products of this tile are in lo_buf
        asm volatile("bar.sync 2, 64;" ::: "memory");     // warp 0 has read them
      } else {
        asm volatile("bar.sync 1, 64;" ::: "memory");
        // is each candidate's pid inside this store?  (lane j < n answers for candidate j)
        int64_t pid = 0;
        if (lane < n) pid = cand_pids[tl.c + lane] - pid_base;
        const uint32_t bad = __ballot_sync(0xffffffffu, lane < n && (pid < 0 || pid >= n_docs));
        float mine = 0.f;               // lane j ends up with the score of candidate j
#pragma unroll 1
        for (int j = 0; j < kMvTileDocs; ++j) {
          uint32_t v[16];
          umma::tmem_ld_32x16(t_addr + j * 16, v);
          umma::tmem_ld_wait();
          float m = -INFINITY;
          if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) m = fmaxf(m, __uint_as_float(v[i]) + lo_buf[(j * 16 + i) * 16 + lane]);
          } else {
            m = 0.f;
          }
          m += __shfl_xor_sync(0xffffffffu, m, 8);
          m += __shfl_xor_sync(0xffffffffu, m, 4);
          m += __shfl_xor_sync(0xffffffffu, m, 2);
          m += __shfl_xor_sync(0xffffffffu, m, 1);
          const float sc = __shfl_sync(0xffffffffu, m, 0);
          if (lane == j) mine = sc;
        }
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[slot]));
        asm volatile("bar.sync 2, 64;" ::: "memory");
        if (lane < n) {
          if ((bad >> lane) & 1u) mine = skip_foreign ? -INFINITY : __int_as_float(0x7fc00000);
          out[tl.c + lane] = mine;
        }
      }
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == 2) umma::tmem_dealloc(tmem, 256);
}

}  // namespace

// the shapes this kernel takes: a fixed-length store of 16 rows per document, queries of at most 16 rows, widths
// 256 … 1024 in steps of 64 (the query region is dim/64 × 8 KB of shared memory; the ring gets what is left: 7 stages
// of 16 KB at 768 columns, 5 at 1024)
bool rerank_mv_wide_supports(int dim, int q_len, const int32_t* strides, int n_strides, int flags) {
  return (flags & CBK_FLAG_FIXED_DOCLEN) && n_strides == 1 && strides[0] == kMvRows && q_len <= kMvRows && dim % 64 == 0 &&
         dim >= 256 && dim <= 1024 && !(flags & (CBK_FLAG_BF16_NATIVE_MMA | CBK_FLAG_RERANK_GENERIC));
}

int rerank_mv_wide_dispatch(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, int64_t n_docs, int64_t pid_base,
                            const float* d_Q, const int32_t* d_q_lens, int q_len, int64_t n_queries, const int64_t* d_cand_pids,
                            const int64_t* d_cand_rowptr, int64_t n_cand_total, float* d_out_scores, int flags,
                            cudaStream_t stream) {
  static thread_local const void* cached_base = nullptr;
  static thread_local int64_t cached_rows = -1;
  static thread_local int cached_dim = -1;
  static thread_local CUtensorMap tmap;
  if (cached_base != d_store || cached_rows != n_store_rows || cached_dim != dim) {
    int rc = make_store_tensor_map(&tmap, d_store, n_store_rows, dim, 64, kMvRows);
    if (rc != CBK_OK) return rc;
    cached_base = d_store;
    cached_rows = n_store_rows;
    cached_dim = dim;
  }
  const int n_slabs = dim / 64;
  const size_t fixed = 1024 + static_cast<size_t>(n_slabs) * kMvASlabBytes + kMvLoBufBytes;
  const int n_stages = static_cast<int>(std::min<size_t>(16, (220 * 1024 - fixed) / kMvStageBytes));
  const size_t smem = fixed + static_cast<size_t>(n_stages) * kMvStageBytes;
  const bool bf16 = store_dtype == CBK_BF16;
  const uint32_t fmt = bf16 ? umma::kFmtBF16 : umma::kFmtF16;
  const uint32_t idesc = umma::make_idesc(64, 128, fmt, fmt);
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sm_count(), (n_cand_total + kMvTileDocs - 1) / kMvTileDocs)));
  const int skip = (flags & CBK_FLAG_SKIP_FOREIGN_PIDS) ? 1 : 0;
  if (bf16) {
    CBK_CUDA(cudaFuncSetAttribute(maxsim_mv_wide_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    maxsim_mv_wide_kernel<__nv_bfloat16><<<grid, kMvThreads, smem, stream>>>(tmap, n_docs, pid_base, skip, d_Q, d_q_lens, q_len, dim,
                                                                           n_queries, d_cand_pids, d_cand_rowptr, n_cand_total,
                                                                           n_stages, idesc, d_out_scores);
  } else {
    CBK_CUDA(cudaFuncSetAttribute(maxsim_mv_wide_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    maxsim_mv_wide_kernel<__half><<<grid, kMvThreads, smem, stream>>>(tmap, n_docs, pid_base, skip, d_Q, d_q_lens, q_len, dim, n_queries,
                                                                    d_cand_pids, d_cand_rowptr, n_cand_total, n_stages, idesc,
                                                                    d_out_scores);
  }
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk
