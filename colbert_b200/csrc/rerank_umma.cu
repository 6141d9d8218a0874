// MaxSim rerank on tcgen05 / TMEM (sm_100a) — second-generation kernel for the same contract as
// maxsim_rerank_kernel (rerank.cu): lookup + exact-length gather by TMA + MaxSim + zero floor, one fp32 per
// candidate written at the candidate's own position.  Replaces reference colbert_ranker.py:88-118 and
// BaseModel.py:41-45.
//
// Orientation: S[tokens, query rows] = Dtile[≤128 tokens, 128] · Q[N, 128]^T  — the document tile is the A
// operand (M = 128 TMEM lanes = tokens), the query is the B operand (N = 32 columns for fp16 stores; for bf16
// stores N = 64: columns 0-31 multiply the bf16 "hi" part of the fp32 query and columns 32-63 its "lo"
// residual — both operands of a tcgen05 MMA must share a format — and the epilogue adds the two halves, which
// keeps 16 significant bits of the query).  Neither operand passes through registers: TMA writes the swizzled
// tile, the tensor core reads it from shared memory, and only the 128 × N accumulator comes back (tcgen05.ld).
//
// One CTA = one autonomous streaming unit (4 per SM for fp16 stores, 3 for bf16):
//   warp 4  producer: claims segments of candidates (the next segment's pid / pfxsum / doclens loads are in
//           flight while the current one is issued), writes the query into one of two shared-memory buffers
//           when it changes, carves variable-size tiles out of a shared-memory ring and issues the TMA loads
//           (exact row count per tile: one tensor map per box height 1..128);
//   warp 5  MMA issuer (one thread): 8 × tcgen05.mma (K = 128) per tile into a ring of TMEM accumulator slots;
//   warps 0-3  epilogue (one warp per TMEM lane quadrant): tcgen05.ld, rows ≥ tile rows masked, max over the 32
//           lanes of every column with one CREDUX (redux.sync.max.f32) per column, the four quadrants combined
//           through shared memory, running max over the chunks of a long document, zero floor, sum over the
//           query rows, store.
// Completion flows back through progress counters (items fully reduced, per epilogue warp; items whose MMAs have
// completed) that the producer and the MMA issuer poll to recycle ring space, item slots and TMEM slots.
//
// Status (round 1): parity with the mma.sync kernel to 1e-6 on fp16 stores and 2e-6 relative error against the
// oracle on bf16 stores, but 15.3-15.5 ms (fp16) / 18.5-18.8 ms (bf16) on configs[1] against 14.4 / 15.4 ms for the mma.sync
// kernel, which therefore stays the default; CBK_FLAG_RERANK_TCGEN05 selects this one.  Measured trade-off: the
// gather is bound by (bytes in flight) / (loaded HBM latency, ≈ 3 µs) and bytes in flight by shared memory.  With
// 4 small CTAs per SM the tile rings total 136 KB (two query buffers and a 128-row-capable ring per CTA eat the
// rest), less than the 192 KB of stages the mma.sync kernel keeps; with 2 big CTAs (176 KB of rings) the single
// producer warp of each CTA cannot issue tiles fast enough (17.3 ms).  Disabling the epilogue's reduction changes
// the time by 3 % only: the tensor-core side is not what limits this kernel.
#include <algorithm>

#include "umma.cuh"

namespace cbk {

namespace {

constexpr int kDim = 128;
constexpr int kTileMax = 128;            // rows per tile (MMA M)
constexpr int kItems = 16;               // item descriptor / full-barrier ring
constexpr int kSegCands = 64;
constexpr int kGroups = 1;               // epilogue groups per CTA (4 warps each)
constexpr int kProducerWarp = kGroups * 4;
constexpr int kMmaWarp = kGroups * 4 + 1;
constexpr int kThreads = (kGroups * 4 + 2) * 32;   // epilogue warps + producer + MMA issuer
constexpr int kTmemCols = 128;           // per CTA (up to four CTAs per SM)

struct StrideSet {
  int n;
  int v[CBK_MAX_STRIDES];
};

struct TileMaps {
  CUtensorMap m[kTileMax];               // m[r-1]: box {64 columns, r rows}
};

struct __align__(16) Item {             // 16 bytes: written with one st.shared.v4, read with one ld.shared.v4
  uint32_t smem_off;                     // tile offset inside the ring
  uint16_t rows;                         // valid rows (1..128); 0 = end of stream
  uint8_t flags;                         // bit0 first chunk, bit1 last chunk, bit2 zero floor, bit3 query buffer, bit4 group
  uint8_t pad;
  uint32_t qseq;                         // query sequence number (the MMA issuer waits for this buffer fill)
  uint32_t out_idx;                      // candidate position
};
static_assert(sizeof(Item) == 16, "Item must stay one 16-byte word");

__device__ __forceinline__ Item load_item(const Item* p) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  Item it;
  *reinterpret_cast<uint4*>(&it) = w;
  return it;
}

struct __align__(16) Shared {
  uint64_t full[kItems];                 // TMA → MMA / epilogue
  uint64_t accf[16];                     // MMA commit → epilogue, per TMEM slot
  uint64_t qfull[2];                     // producer query fill → MMA
  uint64_t ring_free[kItems];            // MMA commit → producer: the tile's shared memory may be overwritten
  Item items[kItems];
  int2 meta[kSegCands];                  // compacted (row, doclen)
  uint8_t cidx[kSegCands];
  float partial[kGroups][2][4][32];      // [group][item parity][quadrant][query row]
  volatile int progress[kGroups * 4];    // per epilogue warp: items (by index) it has looked at and finished with
  uint32_t tmem_base;
};

// every item with index < the returned value has been completely reduced (its ring space, item slot and TMEM
// slot may be reused)
__device__ __forceinline__ int ld_progress_min(const Shared* sh) {
  int m = sh->progress[0];
#pragma unroll
  for (int i = 1; i < kGroups * 4; ++i) m = min(m, sh->progress[i]);
  return m;
}

template <typename T, int kN>
__global__ void __launch_bounds__(kThreads, kN == 32 ? 4 : 3)
maxsim_rerank_umma_kernel(const __grid_constant__ TileMaps maps, const int64_t* __restrict__ pfxsum,
                          const int32_t* __restrict__ doclens, int64_t n_docs, int64_t pid_base, int skip_foreign,
                          StrideSet strides, const float* __restrict__ Q, const int32_t* __restrict__ q_lens, int q_len, int64_t n_queries,
                          const int64_t* __restrict__ cand_pids, const int64_t* __restrict__ rowptr,
                          int64_t n_cand_bound, int seg_cands, int ring_bytes, uint32_t idesc, float* __restrict__ out,
                          unsigned int* __restrict__ seg_counter) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Shared sh;
  constexpr int kSlots = kTmemCols / kN;
  constexpr int kQBufBytes = kN * 256;                               // one query buffer: kN rows × 256 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t ring_addr = (raw + 1023u) & ~1023u;                 // tile ring, then the two query buffers
  const uint32_t q_addr = ring_addr + static_cast<uint32_t>(ring_bytes);
  uint8_t* const q_ptr = smem_raw + (q_addr - raw);

  if (tid == 0) {
    for (int i = 0; i < kItems; ++i) mbar_init(smem_u32(&sh.full[i]), 1);
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&sh.accf[i]), 1);
    for (int i = 0; i < kItems; ++i) mbar_init(smem_u32(&sh.ring_free[i]), 1);
    mbar_init(smem_u32(&sh.qfull[0]), 1);
    mbar_init(smem_u32(&sh.qfull[1]), 1);
    for (int i = 0; i < kGroups * 4; ++i) sh.progress[i] = 0;
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    umma::tmem_alloc(smem_u32(&sh.tmem_base), kTmemCols);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sh.tmem_base;
  const int64_t n_cand = min(rowptr[n_queries], n_cand_bound);
  const int64_t n_segs = (n_cand + seg_cands - 1) / seg_cands;

  // The item stream ends with a sentinel item (rows == 0) that every consumer recognises.
  if (warp == kProducerWarp) {
    // ======================================= producer ===============================================
    for (int r = lane; r < kTileMax; r += 32) tma_prefetch_desc(&maps.m[r]);
    int n_items = 0;             // items published so far
    int head = 0;                // ring allocation cursor
    int live_lo = 0;             // oldest item whose ring space is still accounted (index)
    int tail = 0;                // ring offset of the oldest live item
    int my_start = 0;            // lane i remembers the ring offset of the live item with index ≡ i (mod 32)
    int64_t cur_q = -1;
    uint32_t qseq = 0;
    int qbuf = 1;
    int last_item_of_buf[2] = {-1, -1};
    int doc_parity = 0;

    int known_prog = 0;          // cached lower bounds of the two completion counters (both only grow)
    auto wait_items_done = [&](int upto) {   // until every item with index < upto is finished
      if (known_prog >= upto) return;
      int p = 0;
      if (lane == 0) {
        while ((p = ld_progress_min(&sh)) < upto) __nanosleep(32);
      }
      known_prog = __shfl_sync(0xffffffffu, p, 0);
    };

    // The next segment's metadata is fetched while the current one is being issued: the claim (atomic) and the
    // pid loads right after the current segment has been compacted, the dependent pfxsum / doclens loads a few
    // documents later, so that neither round trip to HBM stalls the TMA stream.
    unsigned int nseg = 0;
    int64_t npid[2] = {-1, -1};
    int nrow[2] = {0, 0}, nlen[2] = {-1, -1};
    auto claim_next = [&]() {
      if (lane == 0) nseg = atomicAdd(seg_counter, 1u);
      nseg = __shfl_sync(0xffffffffu, nseg, 0);
      if (static_cast<int64_t>(nseg) < n_segs) {
        const int64_t b = static_cast<int64_t>(nseg) * seg_cands;
        const int n = static_cast<int>(min(static_cast<int64_t>(seg_cands), n_cand - b));
#pragma unroll
        for (int k = 0; k < 2; ++k) npid[k] = lane + 32 * k < n ? cand_pids[b + lane + 32 * k] - pid_base : -1;
      }
    };
    auto load_next_rows = [&]() {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        nrow[k] = 0;
        nlen[k] = -1;
        if (npid[k] >= 0 && npid[k] < n_docs) {
          nrow[k] = static_cast<int>(pfxsum[npid[k]]);
          nlen[k] = doclens[npid[k]];
        }
      }
    };
    claim_next();
    load_next_rows();

    while (static_cast<int64_t>(nseg) < n_segs) {
      const int64_t c0 = static_cast<int64_t>(nseg) * seg_cands;
      const int nc = static_cast<int>(min(static_cast<int64_t>(seg_cands), n_cand - c0));
      // ---- trivial candidates are answered here, the rest compacted in order ---------------------------
      __syncwarp();
      int nv = 0;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = lane + 32 * k;
        const int len = i < nc ? nlen[k] : -1;
        if (i < nc && len <= 0) out[c0 + i] = len == 0 ? 0.f : (skip_foreign ? -INFINITY : __int_as_float(0x7fc00000));
        const unsigned int live = __ballot_sync(0xffffffffu, len > 0);
        if (len > 0) {
          const int slot = nv + __popc(live & ((1u << lane) - 1u));
          sh.meta[slot] = make_int2(nrow[k], len);
          sh.cidx[slot] = static_cast<uint8_t>(i);
        }
        nv += __popc(live);
      }
      __syncwarp();                                              // the compacted list is visible to every lane
      claim_next();
      const int load_at = min(6, nv - 1);
      if (nv == 0) {
        load_next_rows();
        continue;
      }
      // ---- query of the first scorable candidate -------------------------------------------------------
      const int64_t cfirst = c0 + sh.cidx[0];
      int64_t q = static_cast<int64_t>((static_cast<double>(cfirst) * n_queries) / static_cast<double>(n_cand));
      q = max(static_cast<int64_t>(0), min(q, n_queries - 1));
      if (!(rowptr[q] <= cfirst && cfirst < rowptr[q + 1])) {
        int64_t lo = 0, hi = n_queries - 1;
        while (lo < hi) {
          const int64_t mid = (lo + hi + 1) >> 1;
          if (rowptr[mid] <= cfirst) lo = mid; else hi = mid - 1;
        }
        q = lo;
      }
      int64_t q_end = rowptr[q + 1];

      for (int ci = 0; ci < nv; ++ci) {
        if (ci == load_at) load_next_rows();
        const int64_t c = c0 + sh.cidx[ci];
        while (c >= q_end) {
          ++q;
          q_end = rowptr[q + 1];
        }
        if (q != cur_q) {
          // ---- new query → the other buffer, once no unfinished item still multiplies with it ----------
          cur_q = q;
          qbuf ^= 1;
          ++qseq;
          wait_items_done(last_item_of_buf[qbuf] + 1);
          const float* Qq = Q + q * static_cast<int64_t>(q_len) * kDim;
          const int ql = q_lens ? min(q_len, q_lens[q]) : q_len;   // rows at or past this query's own length read as zero
          uint8_t* dstb = q_ptr + qbuf * kQBufBytes;
          // element (row r, column k) → half h = k / 64, 16-byte chunk (k % 64) / 8 XOR (r & 7) (SWIZZLE_128B);
          // 4 chunks (8 float4 loads) per lane are in flight at a time
          constexpr int kChunks = kN * (kDim / 8);
#pragma unroll 1
          for (int e0 = 0; e0 < kChunks; e0 += 128) {
            float4 f[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e0 + u * 32 + lane;
              const int r = e / (kDim / 8), ch = e % (kDim / 8);
              const int qr = r & 31;                             // rows 32-63 (bf16 store): lo part of row r-32
              f[u][0] = f[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (qr < ql) {
                const float4* src = reinterpret_cast<const float4*>(Qq + qr * kDim + ch * 8);
                f[u][0] = src[0];
                f[u][1] = src[1];
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e0 + u * 32 + lane;
              const int r = e / (kDim / 8), ch = e % (kDim / 8);
              const int h = ch >> 3, cc = ch & 7;
              float v[8] = {f[u][0].x, f[u][0].y, f[u][0].z, f[u][0].w, f[u][1].x, f[u][1].y, f[u][1].z, f[u][1].w};
              uint32_t w[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float a = v[2 * j], b = v[2 * j + 1];
                if (kN == 64 && r >= 32) {                       // residual after rounding to T
                  a -= to_float<T>(static_cast<T>(a));
                  b -= to_float<T>(static_cast<T>(b));
                }
                w[j] = pack2<T>(a, b);
              }
              *reinterpret_cast<uint4*>(dstb + h * (kN * 128) + r * 128 + ((cc ^ (r & 7)) << 4)) =
                  make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sh.qfull[qbuf]));
        }

        const int2 m = sh.meta[ci];
        const int len = m.y;
        bool do_floor = strides.n > 0;
#pragma unroll
        for (int i = 0; i < CBK_MAX_STRIDES; ++i)
          if (i < strides.n && strides.v[i] == len) do_floor = false;
        const int group = doc_parity;
        doc_parity = (doc_parity + 1) % kGroups;
        const int n_chunks = (len + kTileMax - 1) / kTileMax;
        for (int ch = 0; ch < n_chunks; ++ch) {
          const int rows = min(kTileMax, len - ch * kTileMax);
          const int bytes = ((rows + 7) & ~7) * 256;
          // ---- item slot + ring space ----------------------------------------------------------------------
          const int idx = n_items;
          wait_items_done(idx - kItems + 1);                     // slot idx % kItems is free again
          // retire everything the epilogue has finished (keeps live_lo within kItems of idx, so a ring_free
          // barrier is never waited on a whole phase late)
          while (live_lo < known_prog && live_lo < idx) {
            ++live_lo;
            tail = live_lo < idx ? __shfl_sync(0xffffffffu, my_start, live_lo & 31) : head;
          }
          int off = -1;
          while (true) {
            if (live_lo == idx) {                                // ring empty
              head = tail = 0;
              off = 0;
              break;
            }
            // head == tail must only ever mean "empty", hence the strict comparisons against tail
            if (head > tail) {
              if (head + bytes <= ring_bytes) { off = head; break; }
              if (bytes < tail) { off = 0; break; }
            } else if (head + bytes < tail) {
              off = head;
              break;
            }
            // no room: wait until the MMAs of the oldest live tile have completed (tcgen05.commit on its
            // ring_free barrier — independent of the epilogue), then retire it
            mbar_wait(smem_u32(&sh.ring_free[live_lo % kItems]), (live_lo / kItems) & 1);
            ++live_lo;
            tail = live_lo < idx ? __shfl_sync(0xffffffffu, my_start, live_lo & 31) : head;
          }
          head = off + bytes;
          if (lane == (idx & 31)) my_start = off;
          if (elect_one()) {
            Item it;
            it.smem_off = static_cast<uint32_t>(off);
            it.rows = static_cast<uint16_t>(rows);
            it.flags = static_cast<uint8_t>((ch == 0 ? 1 : 0) | (ch == n_chunks - 1 ? 2 : 0) | (do_floor ? 4 : 0) |
                                            (qbuf << 3) | (group << 4));
            it.pad = 0;
            it.qseq = qseq;
            it.out_idx = static_cast<uint32_t>(c);
            *reinterpret_cast<uint4*>(&sh.items[idx % kItems]) = *reinterpret_cast<const uint4*>(&it);
            const uint32_t bar = smem_u32(&sh.full[idx % kItems]);
            const uint32_t dst = ring_addr + off;
            const int row = m.x + ch * kTileMax;
            const CUtensorMap* tm = &maps.m[rows - 1];
            mbar_arrive_expect_tx(bar, rows * 256);
            tma_load_2d(dst, tm, 0, row, bar, kEvictFirst);
            tma_load_2d(dst + ((rows + 7) & ~7) * 128, tm, 64, row, bar, kEvictFirst);
          }
          last_item_of_buf[qbuf] = idx;
          ++n_items;
        }
      }
    }
    // sentinel: tells the MMA issuer and both epilogue groups that the stream is over
    wait_items_done(n_items - kItems + 1);
    if (lane == 0) {
      *reinterpret_cast<uint4*>(&sh.items[n_items % kItems]) = make_uint4(0u, 0u, 0u, 0u);
      mbar_arrive(smem_u32(&sh.full[n_items % kItems]));
    }
  } else if (warp == kMmaWarp) {
    // ======================================= MMA issuer =============================================
    // The whole warp walks the item stream with uniform control flow and one elected lane issues: from a
    // `lane == 0` branch ptxas wraps every tcgen05.mma in an ELECT / R2UR / branch loop (see umma::elect_one).
    uint32_t seen_qseq = 0;
    uint32_t qparity = 0;                                      // bit b = parity of the next wait on qfull[b]
    int known_prog = 0;
    for (int idx = 0;; ++idx) {
      mbar_wait(smem_u32(&sh.full[idx % kItems]), (idx / kItems) & 1);
      const Item it = load_item(&sh.items[idx % kItems]);
      if (it.rows == 0) break;
      const uint32_t qb = (it.flags >> 3) & 1u;
      if (it.qseq != seen_qseq) {                              // first item of a newly written query buffer
        mbar_wait(smem_u32(&sh.qfull[qb]), (qparity >> qb) & 1u);
        qparity ^= 1u << qb;
        seen_qseq = it.qseq;
      }
      while (known_prog < idx - kSlots + 1) {                  // TMEM slot idx % kSlots has been drained
        int p = 0;
        if (lane == 0) p = ld_progress_min(&sh);
        known_prog = __shfl_sync(0xffffffffu, p, 0);
        if (known_prog < idx - kSlots + 1) __nanosleep(32);
      }
      umma::fence_after_sync();
      // descriptors differ only in their 14-bit start-address field: +2 (32 bytes >> 4) per k-step
      const uint32_t a_base = ring_addr + it.smem_off;
      const uint32_t a_half = ((it.rows + 7) & ~7) * 128;
      const uint32_t b_base = q_addr + qb * kQBufBytes;
      const uint64_t a0 = umma::make_smem_desc_sw128(a_base), a1 = umma::make_smem_desc_sw128(a_base + a_half);
      const uint64_t b0 = umma::make_smem_desc_sw128(b_base), b1 = umma::make_smem_desc_sw128(b_base + kN * 128);
      const uint32_t d_tmem = tmem + (idx % kSlots) * kN;
      if (umma::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ss(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ss(d_tmem, a1 + 2 * k, b1 + 2 * k, idesc, 1u);
        umma::commit(smem_u32(&sh.accf[idx % kSlots]));
        umma::commit(smem_u32(&sh.ring_free[idx % kItems]));
      }
      __syncwarp();
    }
  } else {
    // ======================================= epilogue groups ========================================
    const int grp = warp >> 2, quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    float run = -INFINITY;
    int par = 0;
    int known_prog = 0;
    for (int idx = 0;; ++idx) {
      mbar_wait(smem_u32(&sh.full[idx % kItems]), (idx / kItems) & 1);
      const Item it = load_item(&sh.items[idx % kItems]);
      if (it.rows == 0) break;
      if (((it.flags >> 4) & 1) != grp) {                        // the other group's document
        __syncwarp();
        if (lane == 0) sh.progress[warp] = idx + 1;
        continue;
      }
      // the accumulator barrier of a slot is reused every kSlots items, possibly by the other group: wait for
      // the previous use to be completely finished before a parity wait can be unambiguous
      while (known_prog < idx - kSlots + 1) {
        known_prog = ld_progress_min(&sh);
        if (known_prog < idx - kSlots + 1) __nanosleep(32);
      }
      mbar_wait(smem_u32(&sh.accf[idx % kSlots]), (idx / kSlots) & 1);
      umma::fence_after_sync();
      const int nvalid = min(32, max(0, static_cast<int>(it.rows) - 32 * quad));
      if (nvalid > 0) {
        uint32_t raw0[32];
        umma::tmem_ld_32x32(tmem + lane_base + (idx % kSlots) * kN, raw0);
        float v[32];
        if (kN == 64) {   // hi + lo halves of the split query
          uint32_t raw1[32];
          umma::tmem_ld_32x32(tmem + lane_base + (idx % kSlots) * kN + 32, raw1);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw0[j]) + __uint_as_float(raw1[j]);
        } else {
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw0[j]);
        }
        if (lane >= nvalid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = -INFINITY;
        }
        // max over the 32 lanes (tokens) of every column: one warp-wide CREDUX per column (redux.sync.max.f32,
        // sm_100a), then a 5-level select so that lane j keeps the maximum of column j (= query row j)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float r;
          asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v[j]));
          v[j] = r;
        }
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
          const bool up = (lane & w) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) v[i] = up ? v[i + w] : v[i];
        }
        sh.partial[grp][par][quad][lane] = v[0];
      }
      umma::fence_before_sync();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      if (quad == 0) {
        const int nq = (it.rows + 31) >> 5;
        float mx = sh.partial[grp][par][0][lane];
        if (nq > 1) mx = fmaxf(mx, sh.partial[grp][par][1][lane]);
        if (nq > 2) mx = fmaxf(mx, sh.partial[grp][par][2][lane]);
        if (nq > 3) mx = fmaxf(mx, sh.partial[grp][par][3][lane]);
        run = (it.flags & 1) ? mx : fmaxf(run, mx);
        if (it.flags & 2) {
          float t = (it.flags & 4) ? fmaxf(run, 0.f) : run;
          if (lane >= q_len) t = 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) out[it.out_idx] = t;
        }
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        sh.progress[warp] = idx + 1;
      }
      par ^= 1;
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

int rerank_umma_dispatch(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                         const int32_t* d_doclens, int64_t n_docs, int64_t pid_base, const int32_t* strides, int n_strides,
                         const float* d_Q, const int32_t* d_q_lens, int q_len, int64_t n_queries, const int64_t* d_cand_pids,
                         const int64_t* d_cand_rowptr, int64_t n_cand_total, float* d_out_scores, void* d_workspace,
                         int flags, cudaStream_t stream) {
  static thread_local TileMaps maps;
  static thread_local const void* cached_base = nullptr;
  static thread_local int64_t cached_rows = -1;
  if (cached_base != d_store || cached_rows != n_store_rows) {
    cached_base = nullptr;
    for (int r = 1; r <= kTileMax; ++r) {
      int rc = make_store_tensor_map(&maps.m[r - 1], d_store, n_store_rows, dim, 64, r);
      if (rc != CBK_OK) return rc;
    }
    cached_base = d_store;
    cached_rows = n_store_rows;
  }
  CBK_CHECK_SUPPORTED(n_cand_total < (1ll << 32), "cbk_maxsim_rerank (tcgen05): more than 2^32 candidates in one launch");
  StrideSet ss;
  ss.n = n_strides;
  for (int i = 0; i < CBK_MAX_STRIDES; ++i) ss.v[i] = i < n_strides ? strides[i] : -1;
  const int skip = (flags & CBK_FLAG_SKIP_FOREIGN_PIDS) ? 1 : 0;
  unsigned int* counter = static_cast<unsigned int*>(d_workspace);
  CBK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
  const bool bf16 = store_dtype == CBK_BF16;
  // per CTA: 1 KB alignment + tile ring + two query buffers; 4 CTAs per SM for fp16 stores (N = 32), 3 for bf16 (N = 64)
  const int ctas_per_sm = bf16 ? 3 : 4;
  const int ring_bytes = bf16 ? 38 * 1024 : 34 * 1024;
  const size_t smem = 1024 + static_cast<size_t>(ring_bytes) + 2 * (bf16 ? 64 : 32) * 256;
  const int64_t ctas_total = static_cast<int64_t>(sm_count()) * ctas_per_sm;
  const int seg_cands = static_cast<int>(std::max<int64_t>(4, std::min<int64_t>(kSegCands, n_cand_total / (2 * ctas_total))));
  const int64_t n_segs = (n_cand_total + seg_cands - 1) / seg_cands;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_segs, ctas_total)));
  if (bf16) {
    auto kern = maxsim_rerank_umma_kernel<__nv_bfloat16, 64>;
    CBK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kThreads, smem, stream>>>(maps, d_pfxsum, d_doclens, n_docs, pid_base, skip, ss, d_Q, d_q_lens, q_len, n_queries,
                                           d_cand_pids, d_cand_rowptr, n_cand_total, seg_cands, ring_bytes,
                                           umma::make_idesc(128, 64, umma::kFmtBF16, umma::kFmtBF16), d_out_scores, counter);
  } else {
    auto kern = maxsim_rerank_umma_kernel<__half, 32>;
    CBK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kThreads, smem, stream>>>(maps, d_pfxsum, d_doclens, n_docs, pid_base, skip, ss, d_Q, d_q_lens, q_len, n_queries,
                                           d_cand_pids, d_cand_rowptr, n_cand_total, seg_cands, ring_bytes,
                                           umma::make_idesc(128, 32, umma::kFmtF16, umma::kFmtF16), d_out_scores, counter);
  }
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk
