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
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer (one thread), warps 2-17 epilogue: four groups of
//     four warps (one TMEM lane quadrant = one query each).  A CTA owns four document-aligned token
//     sub-ranges and interleaves their tiles; group g drains the accumulators of sub-range g, so four
//     accumulators are being reduced while the next ones are being multiplied.  3-stage smem ring for
//     document tiles, 4 TMEM accumulator slots of 128 columns.
// bf16 stores: both MMA operands must share a format (mixed fp16 × bf16 is an illegal instruction),
// so the query is split into bf16 hi + lo parts and each tile is multiplied twice into the same
// accumulator (K = 256); fp16 stores need one pass over K.
#include <algorithm>

#include "umma.cuh"

namespace cbk {

namespace {

constexpr int kTileTok = 128;
constexpr int kBStages = 3;
constexpr int kABlocks = 4;               // 32 KB A slots in shared memory
constexpr int kAccSlots = 4;              // × 128 TMEM columns
constexpr int kTileBytes = kTileTok * 256;
constexpr int kEpiGroups = 4;             // epilogue warp groups (4 warps each, one per TMEM lane quadrant)
constexpr int kExhThreads = 64 + kEpiGroups * 128;

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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// End of a document reached by a whole warp (lane = query row): apply the reference's zero floor
// (doclen ∉ strides, SURVEY.md §8 a12'), sum the 32 rows, store one score.  Deliberately NOT inlined:
// it is reached from 32 unrolled column positions and the epilogue must stay inside the instruction cache.
__device__ __noinline__ void finish_document(float row_max, int doclen, const int* s_strides, int n_strides,
                                             float* __restrict__ dst, bool write) {
  bool do_floor = n_strides > 0;
  for (int i = 0; i < n_strides; ++i)
    if (s_strides[i] == doclen) do_floor = false;
  const float total = warp_sum(do_floor ? fmaxf(row_max, 0.f) : row_max);
  if (write && (threadIdx.x & 31) == 0) *dst = total;
}

// =====================================================================================================
__global__ void __launch_bounds__(kExhThreads, 1)
maxsim_exhaustive_kernel(const __grid_constant__ ExhMaps maps, const uint32_t* __restrict__ doc_end_bits,
                         const int64_t* __restrict__ pfxsum, const int64_t* __restrict__ cta_doc_start,
                         StrideSet strides, int n_queries, int n_qblocks, int parts, int64_t n_docs, uint32_t idesc,
                         float* __restrict__ scores) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a_full, bar_pass_done;
  __shared__ __align__(8) uint64_t bar_b_full[kBStages], bar_b_empty[kBStages];
  // acc_full is per (consuming epilogue group, slot): an mbarrier wait only tells phases apart by parity, so a
  // barrier must never have two waiters that are a whole phase apart (two groups sharing one slot would be)
  __shared__ __align__(8) uint64_t bar_acc_full[kEpiGroups * kAccSlots], bar_acc_empty[kAccSlots];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int s_strides[CBK_MAX_STRIDES];
  __shared__ int s_n_strides;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < CBK_MAX_STRIDES) s_strides[tid] = strides.v[tid];
  if (tid == 0) s_n_strides = strides.n;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_addr = (raw + 1023u) & ~1023u;              // kABlocks × 32 KB
  const uint32_t b_addr = a_addr + kABlocks * kTileBytes;      // kBStages × 32 KB

  if (tid == 0) {
    mbar_init(smem_u32(&bar_a_full), 1);
    mbar_init(smem_u32(&bar_pass_done), 1);
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(smem_u32(&bar_b_full[s]), 1);
      mbar_init(smem_u32(&bar_b_empty[s]), 1);
    }
    for (int s = 0; s < kEpiGroups * kAccSlots; ++s) mbar_init(smem_u32(&bar_acc_full[s]), 1);
    for (int s = 0; s < kAccSlots; ++s) mbar_init(smem_u32(&bar_acc_empty[s]), 4);   // one arrival per epilogue warp
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

  // four document-aligned sub-ranges per CTA; item (t, g) = tile t of sub-range g, visited t-major
  int64_t sub_d0[kEpiGroups], sub_tok0[kEpiGroups], sub_tok1[kEpiGroups];
  int sub_nt[kEpiGroups];
  int max_nt = 0;
#pragma unroll
  for (int g = 0; g < kEpiGroups; ++g) {
    sub_d0[g] = cta_doc_start[blockIdx.x * kEpiGroups + g];
    const int64_t dend = cta_doc_start[blockIdx.x * kEpiGroups + g + 1];
    sub_tok0[g] = pfxsum[sub_d0[g]];
    sub_tok1[g] = pfxsum[dend];
    sub_nt[g] = static_cast<int>((sub_tok1[g] - sub_tok0[g] + kTileTok - 1) / kTileTok);
    max_nt = max(max_nt, sub_nt[g]);
  }
  const int qb_max = kABlocks / parts;                          // query blocks per pass
  const int n_passes = (n_qblocks + qb_max - 1) / qb_max;

  if (warp == 0) {
    // ===================================== TMA producer =============================================
    if (lane == 0) {
      tma_prefetch_desc(&maps.store);
      tma_prefetch_desc(&maps.q);
      uint32_t it = 0;
      for (int p = 0; p < n_passes; ++p) {
        if (p > 0) mbar_wait(smem_u32(&bar_pass_done), (p - 1) & 1);   // every MMA that read the old A blocks is done
        const int qb = min(qb_max, n_qblocks - p * qb_max);
        const uint32_t afull = smem_u32(&bar_a_full);
        mbar_arrive_expect_tx(afull, static_cast<uint32_t>(qb * parts) * kTileBytes);
        for (int a = 0; a < qb; ++a)
          for (int part = 0; part < parts; ++part) {
            const int row = (part * n_qblocks + p * qb_max + a) * 128;
            const uint32_t dst = a_addr + (a * parts + part) * kTileBytes;
            tma_load_2d(dst, &maps.q, 0, row, afull, kEvictLast);
            tma_load_2d(dst + kTileBytes / 2, &maps.q, 64, row, afull, kEvictLast);
          }
        for (int t = 0; t < max_nt; ++t)
#pragma unroll
          for (int g = 0; g < kEpiGroups; ++g) {
            if (t >= sub_nt[g]) continue;
            const uint32_t st = it % kBStages;
            mbar_wait(smem_u32(&bar_b_empty[st]), ((it / kBStages) & 1u) ^ 1u);
            const uint32_t full = smem_u32(&bar_b_full[st]);
            const uint32_t dst = b_addr + st * kTileBytes;
            const int row = static_cast<int>(sub_tok0[g]) + t * kTileTok;
            mbar_arrive_expect_tx(full, kTileBytes);
            tma_load_2d(dst, &maps.store, 0, row, full, kEvictFirst);
            tma_load_2d(dst + kTileBytes / 2, &maps.store, 64, row, full, kEvictFirst);
            ++it;
          }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ===============================================
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (int p = 0; p < n_passes; ++p) {
        const int qb = min(qb_max, n_qblocks - p * qb_max);
        mbar_wait(smem_u32(&bar_a_full), p & 1);
        umma::fence_after_sync();
        for (int t = 0; t < max_nt; ++t)
#pragma unroll
          for (int g = 0; g < kEpiGroups; ++g) {
            if (t >= sub_nt[g]) continue;
            const uint32_t st = it % kBStages;
            mbar_wait(smem_u32(&bar_b_full[st]), (it / kBStages) & 1u);
            umma::fence_after_sync();
            const uint32_t bt = b_addr + st * kTileBytes;
            for (int a = 0; a < qb; ++a, ++acc_it) {
              const uint32_t slot = acc_it % kAccSlots;
              mbar_wait(smem_u32(&bar_acc_empty[slot]), ((acc_it / kAccSlots) & 1u) ^ 1u);
              umma::fence_after_sync();
              const uint32_t d_tmem = tmem + slot * kTileTok;
              uint32_t acc = 0;
              for (int part = 0; part < parts; ++part) {
                const uint32_t at = a_addr + (a * parts + part) * kTileBytes;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    umma::mma_f16_ss(d_tmem, umma::make_smem_desc_sw128(at + h * (kTileBytes / 2) + k * 32),
                                     umma::make_smem_desc_sw128(bt + h * (kTileBytes / 2) + k * 32), idesc, acc);
                    acc = 1;
                  }
              }
              umma::commit(smem_u32(&bar_acc_full[g * kAccSlots + slot]));
            }
            umma::commit(smem_u32(&bar_b_empty[st]));
            ++it;
          }
        umma::commit(smem_u32(&bar_pass_done));
      }
    }
  } else {
    // ===================================== epilogue (warps 2..17) ===================================
    // Code-size discipline: the loops over query blocks (a) and 32-column chunks (c) are real loops;
    // only the columns of a chunk are unrolled (register array), and the per-document tail is a call.
    const int grp = (warp - 2) >> 2;                 // sub-range this warp group drains
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read = query within the block
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    int64_t my_d0 = 0, my_tok0 = 0, my_tok1 = 0;
#pragma unroll
    for (int g = 0; g < kEpiGroups; ++g)
      if (g == grp) {
        my_d0 = sub_d0[g];
        my_tok0 = sub_tok0[g];
        my_tok1 = sub_tok1[g];
      }
    uint32_t acc_it = 0;
    uint32_t full_parity = 0;                        // bit s = parity of this group's next wait on its barrier of slot s
    for (int p = 0; p < n_passes; ++p) {
      const int qb = min(qb_max, n_qblocks - p * qb_max);
      int64_t doc = my_d0;
      int cur_len = 0;
      float run0 = -INFINITY, run1 = -INFINITY, run2 = -INFINITY, run3 = -INFINITY;

      // document-end bits of a tile: 5 words starting at the tile's first row (prefetched one tile ahead)
      uint32_t wraw[5];
      auto load_end_bits = [&](int t) {
        const int64_t w0 = (my_tok0 + static_cast<int64_t>(t) * kTileTok) >> 5;
#pragma unroll
        for (int i = 0; i < 5; ++i) wraw[i] = doc_end_bits[w0 + i];
      };
      load_end_bits(0);

      for (int t = 0; t < max_nt; ++t) {
        // accumulator slots are handed out in item order: skip over the other groups' items of this round
        bool mine = false;
#pragma unroll
        for (int g = 0; g < kEpiGroups; ++g) {
          if (t >= sub_nt[g]) continue;
          if (g < grp) acc_it += qb;
          if (g == grp) mine = true;
        }
        uint32_t after = 0;
#pragma unroll
        for (int g = 0; g < kEpiGroups; ++g)
          if (t < sub_nt[g] && g > grp) after += qb;
        if (!mine) {
          acc_it += after;
          continue;
        }
        // shift so that bit j of word c is column 32c + j; drop the rows past this sub-range
        const int64_t tbase = my_tok0 + static_cast<int64_t>(t) * kTileTok;
        const int sh = static_cast<int>(tbase & 31);
        uint32_t ends[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ends[c] = __funnelshift_r(wraw[c], wraw[c + 1], sh);
          const int64_t left = my_tok1 - (tbase + 32 * c);
          if (left <= 0) ends[c] = 0u;
          else if (left < 32) ends[c] &= (1u << left) - 1u;
        }
        load_end_bits(t + 1);   // in flight while this tile is reduced (the bitmap is padded past the store)

        int64_t doc_next = doc;
        int len_next = cur_len;
#pragma unroll 1
        for (int a = 0; a < qb; ++a) {
          const uint32_t slot = acc_it % kAccSlots;
          mbar_wait(smem_u32(&bar_acc_full[grp * kAccSlots + slot]), (full_parity >> slot) & 1u);
          full_parity ^= 1u << slot;
          umma::fence_after_sync();
          ++acc_it;
          const int q = (p * qb_max + a) * 4 + quad;
          float* const dst_row = scores + static_cast<int64_t>(q < n_queries ? q : 0) * n_docs;
          int64_t doc_a = doc;
          int len_a = cur_len;
          float r = a == 0 ? run0 : (a == 1 ? run1 : (a == 2 ? run2 : run3));
          const uint32_t t_addr = tmem + lane_base + slot * kTileTok;

          // reduce one 32-column chunk held in registers; 8 columns at a time (most groups of 8 hold no document end)
          auto reduce_chunk = [&](const uint32_t (&v)[32], uint32_t m) {
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
              const uint32_t m8 = (m >> (8 * s8)) & 0xffu;
              if (m8 == 0u) {
                const float x0 = fmaxf(fmaxf(__uint_as_float(v[8 * s8]), __uint_as_float(v[8 * s8 + 1])), __uint_as_float(v[8 * s8 + 2]));
                const float x1 = fmaxf(fmaxf(__uint_as_float(v[8 * s8 + 3]), __uint_as_float(v[8 * s8 + 4])), __uint_as_float(v[8 * s8 + 5]));
                const float x2 = fmaxf(fmaxf(__uint_as_float(v[8 * s8 + 6]), __uint_as_float(v[8 * s8 + 7])), r);
                r = fmaxf(fmaxf(x0, x1), x2);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  r = fmaxf(r, __uint_as_float(v[8 * s8 + j]));
                  if ((m8 >> j) & 1u) {   // this column is the last token of document doc_a (warp-uniform)
                    finish_document(r, len_a + 8 * s8 + j + 1, s_strides, s_n_strides, dst_row + doc_a, q < n_queries);
                    ++doc_a;
                    len_a = -(8 * s8 + j + 1);
                    r = -INFINITY;
                  }
                }
              }
            }
            len_a += 32;
          };

          // TMEM loads are double-buffered in registers: chunk c+1 is in flight while chunk c is reduced
          uint32_t va[32], vb[32];
          umma::tmem_ld_32x32(t_addr, va);
          umma::tmem_ld_wait();
          umma::tmem_ld_32x32(t_addr + 32, vb);
          reduce_chunk(va, ends[0]);
          umma::tmem_ld_wait();
          umma::tmem_ld_32x32(t_addr + 64, va);
          reduce_chunk(vb, ends[1]);
          umma::tmem_ld_wait();
          umma::tmem_ld_32x32(t_addr + 96, vb);
          reduce_chunk(va, ends[2]);
          umma::tmem_ld_wait();
          // all columns of the slot are in registers: hand it back to the MMA warp
          umma::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[slot]));
          reduce_chunk(vb, ends[3]);

          if (a == 0) run0 = r; else if (a == 1) run1 = r; else if (a == 2) run2 = r; else run3 = r;
          doc_next = doc_a;
          len_next = len_a;
        }
        doc = doc_next;
        cur_len = len_next;
        acc_it += after;
      }
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

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
  const size_t smem = 1024 + static_cast<size_t>(kABlocks + kBStages) * kTileBytes;
  CBK_CUDA(cudaFuncSetAttribute(maxsim_exhaustive_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  maxsim_exhaustive_kernel<<<n_ctas, kExhThreads, smem, stream>>>(maps, d_doc_end_bits, d_pfxsum, d_ranges, ss,
                                                                static_cast<int>(n_queries), n_qblocks, parts, n_docs, idesc,
                                                                d_out_scores);
  CBK_CUDA(cudaGetLastError());
  count_launch(3);
  return CBK_OK;
}

}  // namespace cbk
