// MaxSim rerank for wide embeddings (dim = 64·k, 192 … 1024; the author's index is 768 wide, reference
// proj_conf/dense.yaml:8) on tcgen05 / TMEM / TMA: documents of any length — including the author's multi-view layout of
// exactly 16 view embeddings per document and per query (dense.yaml:29-32, BaseModel.get_representation BaseModel.py:21-27),
// for which CBK_FLAG_FIXED_DOCLEN skips the pfxsum / doclens lookups — queries of up to 32 rows, zero-floor rule.  Same
// contract as cbk_maxsim_rerank (reference colbert_ranker.py:88-126, BaseModel.py:39-46).
//
// A row is up to 2 KB, the path is HBM-bound (≤ 32 FLOP/B).  The K-split mma.sync kernel (rerank_wide.cu) pays a block
// barrier and an 8-way shared-memory reduction per 16-row tile (0.69 of the copy peak inside bench.py at 768); this kernel
// streams:
//   * every candidate document is cut into PIECES of up to 16 rows (the last one holds the remainder); a tile is up to 16
//     consecutive pieces of candidates of ONE query — piece j occupies accumulator columns 16 j … 16 j + 15 — and one 64-column
//     K slab of a tile is a ring stage (32 KB): one TMA box per piece, fetched with a tensor map of exactly the piece's height
//     (one map per height 1 … 16), so no byte outside the candidate documents is read;
//   * the query is resident in shared memory while the CTA stays on it (dim/64 slabs × 64 rows × 128 B), written by a
//     dedicated warp from the fp32 query in the K-major 128-byte-swizzled layout: rows 0–31 = the query rounded to the
//     store's type, rows 32–63 = for bf16 stores the bf16 RESIDUAL of that rounding (zero for fp16 stores), so a bf16 store
//     is multiplied with 16 significant bits of the query at no extra MMA;
//   * tcgen05.mma M = 64, N = 256 into one of two 256-column TMEM accumulators (N = 256 halves the MMAs issued per byte:
//     one issuing warp sustains an MMA every ~88 cycles, and with N = 128 that, not HBM, set the pace); accumulator row m lives in TMEM lane
//     (m mod 16) + 32·(m div 16): quadrant w holds query rows 16 w … (w = 0, 1) and their residual products (w = 2, 3);
//   * epilogue: FOUR teams of up to four warps (one per TMEM quadrant); team k owns the documents whose running number is
//     k mod 4, so a document's running maximum never leaves one warp while four documents are folded at a time (with a single
//     team the epilogue, not HBM, set the pace below dim 768: 0.58 of the copy peak at dim 256, 0.96 with the arithmetic
//     switched off).  In a team, the warp of quadrant w + 2 passes the residual products of a piece to the warp of quadrant w
//     through a two-slot shared-memory buffer guarded by mbarriers; that warp folds the piece's valid columns into the running
//     maximum per query row and, at a document's last piece, applies the zero floor (doclen ∉ strides, SURVEY.md §8 a12'),
//     adds up its 16 rows and adds the sum to the score at the candidate's own position (two commutative additions into a
//     zeroed score when the query has more than 16 rows: bit-reproducible).
// Warp roles: 0–15 epilogue (quadrant = warp mod 4, team = warp div 4) · 16 MMA issuer · 17 query loader · 18–21 TMA
// producers · 22 planner.  The planner walks the CTA's
// contiguous candidate range (CandWalker: 32 candidates' metadata fetched at a time, one per lane, the next 32 prefetched),
// cuts it into tiles and publishes one descriptor per tile in a shared-memory ring; every other role reads descriptors —
// walking the candidates in each role cost the issuer ~2000 cycles per tile on its critical path.
#include <algorithm>
#include <cstdlib>

#include "umma.cuh"

namespace cbk {

int make_store_tensor_map(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_cols, int box_rows);

namespace {

constexpr int kWsPieceRows = 16;
constexpr int kWsTilePieces = 16;
constexpr int kWsSlotBytes = kWsPieceRows * 128;                  // a piece's slot in a stage (2 KB)
constexpr int kWsStageBytes = kWsTilePieces * kWsSlotBytes;       // 32 KB
constexpr int kWsTileCols = kWsTilePieces * kWsPieceRows;         // 256 accumulator columns
constexpr int kWsASlabBytes = 64 * 128;
constexpr int kWsProducers = 4;
constexpr int kWsTeams = 4;                                       // epilogue teams: documents are dealt round-robin to them
constexpr int kWsEpiWarps = 4 * kWsTeams;
constexpr int kWsIssuerWarp = kWsEpiWarps, kWsLoaderWarp = kWsEpiWarps + 1, kWsProducer0 = kWsEpiWarps + 2,
              kWsPlannerWarp = kWsProducer0 + kWsProducers;
constexpr int kWsThreads = (kWsPlannerWarp + 1) * 32;
constexpr int kWsLoBufBytes = 2 * kWsTeams * 2 * 16 * 16 * 4;     // residual products: [query-row half][team][slot][column][row] fp32
constexpr int kWsMaxStages = 6;
constexpr int kWsDescRing = 8;

// What the planner publishes per tile (n == 0: the range is exhausted).
struct WsDesc {
  int n;
  int pad;
  int64_t q;
  int row[kWsTilePieces];
  int len[kWsTilePieces];
  int64_t cand[kWsTilePieces];
  uint8_t v[kWsTilePieces];
  uint8_t flags[kWsTilePieces];       // bit 0 first piece, bit 1 last piece of its document, bit 2 the zero floor applies,
                                      // bits 3-4 the epilogue team that owns the document
};

struct WsMaps {
  CUtensorMap m[kWsPieceRows];      // box {64 columns, h rows}, h = 1 … 16
};

struct WsStrides {
  int n;
  int v[CBK_MAX_STRIDES];
};

// One lane's view of the piece it holds in the current tile (lane j < n holds piece j).
struct WsPiece {
  int row;        // first store row of the piece
  int v;          // valid rows (1 … 16)
  int flags;      // bit 0: first piece of its document, bit 1: last piece
  int len;        // length of the document (floor rule)
  int64_t cand;   // position of the candidate (where its score goes)
};

// Walks the candidates [c, c_hi) of this CTA as pieces grouped in tiles.  Every field is warp-uniform except the window
// (win_*: lane l holds candidate win_c0 + l) — all 32 lanes of a role's warp call next_tile() together.
struct CandWalker {
  const int64_t* cand_pids;
  const int64_t* rowptr;
  const int64_t* pfxsum;
  const int32_t* doclens;
  int64_t n_docs, pid_base, c_hi;
  int fixed_len;                    // > 0: every document has exactly this many rows (CBK_FLAG_FIXED_DOCLEN): no metadata is read
  int64_t c_fetch;                  // first candidate of the window AFTER the prefetched one
  int64_t win_c0;
  int win_n, pos;
  int win_row, win_len, nxt_row, nxt_len, nxt_n;        // lane-private: current and prefetched window
  int64_t nxt_c0;
  int64_t q, q_end;                 // query of the candidate being cut
  int64_t cur_c, cur_q;
  int cur_row, cur_len, cur_left;   // cur_left > 0: a document is being cut
  int doc_seq;                      // documents cut so far (the open one included): deals them to the epilogue teams

  __device__ __forceinline__ void fetch(int lane, int64_t c0, int& row, int& len, int& n) {
    n = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(32), c_hi - c0)));
    row = 0;
    len = -1;
    if (lane < n) {
      const int64_t p = cand_pids[c0 + lane] - pid_base;
      if (p >= 0 && p < n_docs) {
        row = fixed_len > 0 ? static_cast<int>(p) * fixed_len : static_cast<int>(pfxsum[p]);
        len = fixed_len > 0 ? fixed_len : doclens[p];
      }
    }
  }
  __device__ __forceinline__ void init(int lane, const int64_t* cp, const int64_t* rp, const int64_t* pf, const int32_t* dl,
                                       int64_t nd, int64_t pb, int fixed, int64_t n_queries, int64_t lo, int64_t hi) {
    cand_pids = cp; rowptr = rp; pfxsum = pf; doclens = dl; n_docs = nd; pid_base = pb; c_hi = hi; fixed_len = fixed;
    int64_t a = 0, b = n_queries - 1;      // last query whose list starts at or before lo
    while (a < b) {
      const int64_t mid = (a + b + 1) >> 1;
      if (rp[mid] <= lo) a = mid; else b = mid - 1;
    }
    q = a;
    q_end = rp[q + 1];
    win_c0 = lo; win_n = 0; pos = 0;
    cur_left = 0; cur_c = -1; cur_q = -1; cur_row = 0; cur_len = 0; doc_seq = 0;
    nxt_c0 = lo;
    fetch(lane, lo, nxt_row, nxt_len, nxt_n);
    c_fetch = lo + 32;
  }

  // Fills `pc` (lane j < n: piece j) → number of pieces in the tile (0: the range is exhausted); tile_q = its query.
  // kEmit: this caller also writes the scores of candidates that have nothing to multiply (empty document → 0, pid outside
  // the store → NaN, or −inf on a shard).
  template <bool kEmit>
  __device__ __forceinline__ int next_tile(int lane, WsPiece& pc, int64_t& tile_q, float* out, int skip_foreign) {
    int n = 0;
    tile_q = -1;
    while (n < kWsTilePieces) {
      if (cur_left <= 0) {
        if (pos >= win_n) {               // next window: take the prefetched one, prefetch its successor
          if (nxt_n <= 0) break;
          win_c0 = nxt_c0; win_n = nxt_n; win_row = nxt_row; win_len = nxt_len; pos = 0;
          nxt_c0 = c_fetch;
          fetch(lane, c_fetch, nxt_row, nxt_len, nxt_n);
          c_fetch += 32;
        }
        const int len = __shfl_sync(0xffffffffu, win_len, pos);
        const int row = __shfl_sync(0xffffffffu, win_row, pos);
        const int64_t c = win_c0 + pos;
        ++pos;
        if (len <= 0) {
          if (kEmit && lane == 0) out[c] = len == 0 ? 0.f : (skip_foreign ? -INFINITY : __int_as_float(0x7fc00000));
          continue;
        }
        while (c >= q_end) {
          ++q;
          q_end = rowptr[q + 1];
        }
        cur_c = c; cur_q = q; cur_row = row; cur_len = len; cur_left = len;
        ++doc_seq;
      }
      if (n > 0 && cur_q != tile_q) break;       // the tile closes at a query boundary; the document stays pending
      tile_q = cur_q;
      const int v = min(kWsPieceRows, cur_left);
      if (lane == n) {
        pc.row = cur_row + (cur_len - cur_left);
        pc.v = v;
        pc.flags = (cur_left == cur_len ? 1 : 0) | (cur_left <= kWsPieceRows ? 2 : 0) | ((doc_seq & (kWsTeams - 1)) << 3);
        pc.len = cur_len;
        pc.cand = cur_c;
      }
      cur_left -= v;
      ++n;
    }
    return n;
  }
};

template <typename T>
__device__ __forceinline__ void ws_split(float x, T& hi, T& lo);
template <>
__device__ __forceinline__ void ws_split<__half>(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(0.f);
}
template <>
__device__ __forceinline__ void ws_split<__nv_bfloat16>(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// kHasLo: bf16 stores — rows 32-63 of the query region carry the rounding residuals and warps 2, 3 take part in the epilogue
template <typename T, bool kHasLo>
__global__ void __launch_bounds__(kWsThreads, 1)
maxsim_wide_stream_kernel(const __grid_constant__ WsMaps maps, const int64_t* __restrict__ pfxsum, const int32_t* __restrict__ doclens,
                          int64_t n_docs, int64_t pid_base, int fixed_len, int skip_foreign, WsStrides strides,
                          const float* __restrict__ Q,
                          const int32_t* __restrict__ q_lens, int q_len, int dim, int64_t n_queries,
                          const int64_t* __restrict__ cand_pids, const int64_t* __restrict__ rowptr, int64_t n_cand_bound,
                          int n_stages, uint32_t idesc, float* __restrict__ out, int probe_stream_only) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kWsMaxStages], bar_empty[kWsMaxStages];
  __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2], bar_a_full, bar_a_free;
  __shared__ __align__(8) uint64_t bar_desc_full[kWsDescRing], bar_desc_empty[kWsDescRing];
  __shared__ __align__(16) WsDesc descs[kWsDescRing];
  __shared__ __align__(8) uint64_t bar_lo_ready[2 * kWsTeams][2], bar_lo_free[2 * kWsTeams][2];   // [half * teams + team][slot]
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_slabs = dim >> 6;
  const bool two_halves = q_len > 16;            // query rows 16-31 exist: quadrant 1 (and 3) take part
  const uint32_t epi_warps = kWsTeams * (two_halves ? 2u : 1u) * (kHasLo ? 2u : 1u);
  const uint32_t a_addr = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_addr = a_addr + static_cast<uint32_t>(n_slabs) * kWsASlabBytes;
  uint8_t* const smem_al = smem_raw + (a_addr - smem_u32(smem_raw));
  float* const lo_buf = reinterpret_cast<float*>(smem_al + static_cast<size_t>(n_slabs) * kWsASlabBytes +
                                                 static_cast<size_t>(n_stages) * kWsStageBytes);

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), epi_warps);
    }
    for (int s = 0; s < kWsDescRing; ++s) {
      mbar_init(smem_u32(&bar_desc_full[s]), 1);
      mbar_init(smem_u32(&bar_desc_empty[s]), kWsProducers + 2 + epi_warps);    // every reading warp releases a descriptor once
    }
    mbar_init(smem_u32(&bar_a_full), 1);
    mbar_init(smem_u32(&bar_a_free), 1);
    for (int i = 0; i < 2 * kWsTeams; ++i)
      for (int sl = 0; sl < 2; ++sl) {
        mbar_init(smem_u32(&bar_lo_ready[i][sl]), 1);
        mbar_init(smem_u32(&bar_lo_free[i][sl]), 1);
      }
    fence_mbar_init();
  }
  if (warp == kWsIssuerWarp) {
    umma::tmem_alloc(smem_u32(&tmem_base_smem), 512);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;

  // descriptor ring, reader side: wait for tile `t`'s descriptor / hand it back once its fields are in registers
  auto desc_wait = [&](uint32_t t) -> const WsDesc* {
    mbar_wait(smem_u32(&bar_desc_full[t % kWsDescRing]), (t / kWsDescRing) & 1u);
    return &descs[t % kWsDescRing];
  };
  auto desc_release = [&](uint32_t t) {
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_desc_empty[t % kWsDescRing]));
  };

  if (warp == kWsPlannerWarp) {
    // ===================================== planner ==================================================
    const int64_t n_cand = min(rowptr[n_queries], n_cand_bound);
    const int64_t per = (n_cand + gridDim.x - 1) / gridDim.x;
    const int64_t c_lo = min(n_cand, per * blockIdx.x), c_hi = min(n_cand, per * (blockIdx.x + 1));
    CandWalker wk;
    wk.init(lane, cand_pids, rowptr, pfxsum, doclens, n_docs, pid_base, fixed_len, n_queries, c_lo, c_hi);
    WsPiece pc;
    pc.row = 0; pc.v = 0; pc.flags = 0; pc.len = 0; pc.cand = 0;
    int64_t tile_q = -1;
    for (uint32_t t = 0;; ++t) {
      const int n = wk.next_tile<true>(lane, pc, tile_q, out, skip_foreign);
      mbar_wait(smem_u32(&bar_desc_empty[t % kWsDescRing]), ((t / kWsDescRing) & 1u) ^ 1u);
      WsDesc* d = &descs[t % kWsDescRing];
      if (lane < n) {
        int fl = pc.flags;
        if (pc.flags & 2) {   // the zero floor of SURVEY.md §8 a12' applies when the document's length is not one of the strides
          bool floor0 = strides.n > 0;
          for (int i = 0; i < strides.n; ++i)
            if (strides.v[i] == pc.len) floor0 = false;
          if (floor0) fl |= 4;
        }
        d->row[lane] = pc.row;
        d->len[lane] = pc.len;
        d->cand[lane] = pc.cand;
        d->v[lane] = static_cast<uint8_t>(pc.v);
        d->flags[lane] = static_cast<uint8_t>(fl);
      }
      if (lane == 0) {
        d->n = n;
        d->q = tile_q;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_desc_full[t % kWsDescRing]));
      if (n == 0) break;
    }
  } else if (warp >= kWsProducer0) {
    // ===================================== TMA producers ============================================
    const int p = warp - kWsProducer0;
    if (lane == 0 && p == 0) tma_prefetch_desc(&maps.m[kWsPieceRows - 1]);
    uint32_t it = 0;
    for (uint32_t t = 0;; ++t) {
      const WsDesc* d = desc_wait(t);
      const int n = d->n;
      int my_row = 0, my_v = 0;                          // lane j < n: piece j
      if (lane < n) {
        my_row = d->row[lane];
        my_v = d->v[lane];
      }
      desc_release(t);
      if (n == 0) break;
      int rows_total = my_v;                            // bytes of a stage = 128 B × the rows of all its pieces
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rows_total += __shfl_xor_sync(0xffffffffu, rows_total, o);
      for (int s = 0; s < n_slabs; ++s, ++it) {
        const uint32_t st = it % static_cast<uint32_t>(n_stages);
        mbar_wait(smem_u32(&bar_empty[st]), ((it / static_cast<uint32_t>(n_stages)) & 1u) ^ 1u);
        const uint32_t full = smem_u32(&bar_full[st]);
        const uint32_t dst = b_addr + st * kWsStageBytes;
        // producer 0 arms the barrier with the bytes of the WHOLE stage; the other producers' boxes may land before that
        // (the transaction count goes negative meanwhile, the phase cannot complete before the arrival)
        if (p == 0) {
          if (elect_one()) mbar_arrive_expect_tx(full, static_cast<uint32_t>(rows_total) * 128u);
          __syncwarp();
        }
        for (int j = p; j < n; j += kWsProducers) {     // my pieces of the tile: one box each, of exactly the piece's height
          const int rj = __shfl_sync(0xffffffffu, my_row, j);
          const int vj = __shfl_sync(0xffffffffu, my_v, j);
          if (elect_one()) tma_load_2d(dst + j * kWsSlotBytes, &maps.m[vj - 1], s * 64, rj, full, kEvictFirst);
        }
        __syncwarp();
      }
    }
  } else if (warp == kWsIssuerWarp) {
    // ===================================== MMA issuer ===============================================
    const uint32_t full0 = hold(smem_u32(&bar_full[0])), empty0 = hold(smem_u32(&bar_empty[0]));
    const uint32_t a_lo0 = hold(umma::desc_lo_sw128(a_addr)), b_lo0 = hold(umma::desc_lo_sw128(b_addr));
    constexpr uint32_t kStageDesc = kWsStageBytes >> 4, kSlabDesc = kWsASlabBytes >> 4;
    uint32_t st = 0, st_parity = 0, n_q_seen = 0;
    int64_t cur_q = -1;
    for (uint32_t t = 0;; ++t) {
      const WsDesc* d = desc_wait(t);
      const int n = d->n;
      const int64_t tile_q = d->q;
      desc_release(t);
      if (n == 0) break;
      if (tile_q != cur_q) {
        // every MMA issued so far read the old query: its completion frees the query region for the loader
        if (cur_q >= 0) {
          if (elect_one()) umma::commit(smem_u32(&bar_a_free));
          __syncwarp();
        }
        cur_q = tile_q;
        mbar_wait(smem_u32(&bar_a_full), n_q_seen & 1u);
        ++n_q_seen;
        umma::fence_after_sync();
      }
      const uint32_t slot = t & 1u;
      mbar_wait(smem_u32(&bar_acc_empty[slot]), ((t >> 1) & 1u) ^ 1u);
      umma::fence_after_sync();
      const uint32_t d_tmem = tmem + slot * kWsTileCols;
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
  } else if (warp == kWsLoaderWarp) {
    // ===================================== query loader =============================================
    uint32_t n_q_seen = 0;
    int64_t cur_q = -1;
    for (uint32_t t = 0;; ++t) {
      const WsDesc* d = desc_wait(t);
      const int n = d->n;
      const int64_t tile_q = d->q;
      desc_release(t);
      if (n == 0) break;
      if (tile_q == cur_q) continue;
      cur_q = tile_q;
      if (n_q_seen > 0) mbar_wait(smem_u32(&bar_a_free), (n_q_seen - 1) & 1u);    // the MMAs of the previous query are done
      const float* Qq = Q + cur_q * static_cast<int64_t>(q_len) * dim;
      const int ql = q_lens ? min(q_len, q_lens[cur_q]) : q_len;
      // one 16-byte chunk (8 columns) per step: (row r, slab s, chunk c16) → row r (rounded value) and row 32 + r (residual);
      // rows at or past the query's own length are written as zeros
      const int n_chunks = 32 * n_slabs * 8;
#pragma unroll 4
      for (int i = lane; i < n_chunks; i += 32) {
        const int r = i / (n_slabs * 8), s = (i / 8) % n_slabs, c16 = i % 8;
        uint4 hi4 = make_uint4(0u, 0u, 0u, 0u), lo4 = hi4;
        if (r < ql) {
          const float4 f0 = *reinterpret_cast<const float4*>(Qq + static_cast<int64_t>(r) * dim + s * 64 + c16 * 8);
          const float4 f1 = *reinterpret_cast<const float4*>(Qq + static_cast<int64_t>(r) * dim + s * 64 + c16 * 8 + 4);
          const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          T h[8], l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ws_split<T>(x[j], h[j], l[j]);
          hi4 = *reinterpret_cast<const uint4*>(h);
          lo4 = *reinterpret_cast<const uint4*>(l);
        }
        // K-major, 128-byte swizzle: the 16-byte chunk index is XORed with the row's low three bits ((32 + r) & 7 == r & 7)
        uint8_t* slab = smem_al + s * kWsASlabBytes;
        *reinterpret_cast<uint4*>(slab + r * 128 + ((c16 ^ (r & 7)) << 4)) = hi4;
        *reinterpret_cast<uint4*>(slab + (32 + r) * 128 + ((c16 ^ (r & 7)) << 4)) = lo4;
      }
      fence_proxy_async();            // generic-proxy writes → visible to the tensor core's (async-proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_a_full));
      ++n_q_seen;
    }
  } else {
    // ===================================== epilogue (warps 0..15) ===================================
    // quadrant qd = warp % 4: qd < 2 holds query rows 16 qd … 16 qd + 15, qd + 2 the residual products of the same rows;
    // team = warp / 4 owns the documents whose running number is team (mod 4)
    const int qd = warp & 3, team = warp >> 2;
    const bool is_lo = qd >= 2;
    const int half = qd & 1;
    const bool active = (half == 0 || two_halves) && (!is_lo || kHasLo);
    if (active) {
      const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
      const int pair = half * kWsTeams + team;
      float* const my_lo = lo_buf + pair * (2 * 16 * 16);          // two slots of [column][row]
      const uint32_t ready0 = smem_u32(&bar_lo_ready[pair][0]), free0 = smem_u32(&bar_lo_free[pair][0]);
      uint32_t n_mine = 0;                                      // my pieces so far: slot = n_mine & 1, phase = n_mine >> 1
      float rmax = -INFINITY;                                   // running maximum of my open document, per query row
      for (uint32_t t = 0;; ++t) {
        const WsDesc* d = desc_wait(t);
        const int n = d->n;
        int my_v = 0, my_flags = 0;                             // lane j < n: piece j
        int64_t my_cand = 0;
        if (lane < n) {
          my_v = d->v[lane];
          my_flags = d->flags[lane];
          my_cand = d->cand[lane];
        }
        desc_release(t);
        if (n == 0) break;
        uint32_t mine = __ballot_sync(0xffffffffu, lane < n && ((my_flags >> 3) & (kWsTeams - 1)) == team);
        const uint32_t slot = t & 1u;
        mbar_wait(smem_u32(&bar_acc_full[slot]), (t >> 1) & 1u);
        umma::fence_after_sync();
        const uint32_t t_addr = tmem + lane_base + slot * kWsTileCols;
        if (!probe_stream_only) {       // (CBK_WS_PROBE=1: what gather + MMA sustain without the epilogue arithmetic)
          while (mine) {
            const int j = __ffs(mine) - 1;
            mine &= mine - 1u;
            uint32_t v[16];
            umma::tmem_ld_32x16(t_addr + j * 16, v);
            umma::tmem_ld_wait();
            const uint32_t sl = n_mine & 1u, ph = (n_mine >> 1) & 1u;
            ++n_mine;
            if (is_lo) {
              // residual products of this piece → slot sl, once the value warp has read what was there two pieces ago
              mbar_wait(free0 + 8 * sl, ph ^ 1u);
              if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) my_lo[(sl * 16 + i) * 16 + lane] = __uint_as_float(v[i]);
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(ready0 + 8 * sl);
              continue;
            }
            const int pv = __shfl_sync(0xffffffffu, my_v, j);
            const int pf = __shfl_sync(0xffffffffu, my_flags, j);
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
            if (kHasLo) {
              mbar_wait(ready0 + 8 * sl, ph);
              if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] += my_lo[(sl * 16 + i) * 16 + lane];
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(free0 + 8 * sl);
            }
            float m;
            if (pv == kWsPieceRows) {     // a full piece (all but the last of a document): plain max tree
              m = fmaxf(fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7]))),
                        fmaxf(fmaxf(fmaxf(x[8], x[9]), fmaxf(x[10], x[11])), fmaxf(fmaxf(x[12], x[13]), fmaxf(x[14], x[15]))));
            } else {
              m = -INFINITY;
#pragma unroll
              for (int i = 0; i < 16; ++i) m = fmaxf(m, i < pv ? x[i] : -INFINITY);
            }
            rmax = (pf & 1) ? m : fmaxf(rmax, m);
            if (pf & 2) {   // the document ends with this piece: floor, then the sum over my rows (lanes 16-31 hold nothing)
              float sum = lane < 16 ? fmaxf(rmax, (pf & 4) ? 0.f : -INFINITY) : 0.f;
              sum += __shfl_xor_sync(0xffffffffu, sum, 8);
              sum += __shfl_xor_sync(0xffffffffu, sum, 4);
              sum += __shfl_xor_sync(0xffffffffu, sum, 2);
              sum += __shfl_xor_sync(0xffffffffu, sum, 1);
              const int64_t cand = __shfl_sync(0xffffffffu, my_cand, j);
              if (lane == 0) {
                if (two_halves) atomicAdd(out + cand, sum);      // two commutative additions into a zeroed score
                else out[cand] = sum;
              }
            }
          }
        }
        // every column I needed has been read: hand the slot back to the MMA warp
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[slot]));
      }
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == kWsIssuerWarp) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

// widths 192 … 1024 in steps of 64 (the query region takes dim/64 × 8 KB of shared memory and the ring what is left: 3 stages
// of 32 KB at 768, 2 at 1024 — still ahead of the K-split kernel there, 0.83 vs 0.69 of the copy peak), queries of at most 32 rows
bool rerank_wide_stream_supports(int dim, int q_len, int flags) {
  // (dim 128 stays with the per-warp mma.sync kernel of rerank.cu: measured there, this structure reaches 0.30 of the copy
  //  peak — 0.75 with the epilogue arithmetic switched off — against 0.97: at 256 B per row the per-piece work of the
  //  planner, the producers and the single epilogue warp pair outweighs the bytes)
  return dim % 64 == 0 && dim >= 192 && dim <= 1024 && q_len <= 32 &&
         !(flags & (CBK_FLAG_BF16_NATIVE_MMA | CBK_FLAG_RERANK_GENERIC | CBK_FLAG_RERANK_KSPLIT));
}

int rerank_wide_stream_dispatch(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                                const int32_t* d_doclens, int64_t n_docs, int64_t pid_base, const int32_t* strides, int n_strides,
                                const float* d_Q, const int32_t* d_q_lens, int q_len, int64_t n_queries, const int64_t* d_cand_pids,
                                const int64_t* d_cand_rowptr, int64_t n_cand_total, float* d_out_scores, int flags,
                                cudaStream_t stream) {
  static thread_local const void* cached_base = nullptr;
  static thread_local int64_t cached_rows = -1;
  static thread_local int cached_dim = -1;
  static thread_local WsMaps maps;
  if (cached_base != d_store || cached_rows != n_store_rows || cached_dim != dim) {
    for (int h = 1; h <= kWsPieceRows; ++h) {
      int rc = make_store_tensor_map(&maps.m[h - 1], d_store, n_store_rows, dim, 64, h);
      if (rc != CBK_OK) return rc;
    }
    cached_base = d_store;
    cached_rows = n_store_rows;
    cached_dim = dim;
  }
  WsStrides ss;
  ss.n = n_strides;
  for (int i = 0; i < CBK_MAX_STRIDES; ++i) ss.v[i] = i < n_strides ? strides[i] : -1;
  const int n_slabs = dim / 64;
  const size_t fixed = 1024 + static_cast<size_t>(n_slabs) * kWsASlabBytes + (store_dtype == CBK_BF16 ? kWsLoBufBytes : 0);
  const int n_stages = static_cast<int>(std::min<size_t>(kWsMaxStages, (220 * 1024 - fixed) / kWsStageBytes));
  const size_t smem = fixed + static_cast<size_t>(n_stages) * kWsStageBytes;
  const bool bf16 = store_dtype == CBK_BF16;
  const uint32_t fmt = bf16 ? umma::kFmtBF16 : umma::kFmtF16;
  const uint32_t idesc = umma::make_idesc(64, kWsTileCols, fmt, fmt);
  // a candidate is at least one piece; ranges shorter than a few tiles are not worth a CTA
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sm_count(), (n_cand_total + 31) / 32)));
  const int skip = (flags & CBK_FLAG_SKIP_FOREIGN_PIDS) ? 1 : 0;
  // CBK_FLAG_FIXED_DOCLEN: the caller guarantees doclens[p] == strides[0] for every document (multi-view index)
  const int fixed_len = ((flags & CBK_FLAG_FIXED_DOCLEN) && n_strides == 1 && strides[0] > 0) ? strides[0] : 0;
  static const int probe = std::getenv("CBK_WS_PROBE") != nullptr;          // profiling aid, never set in production
  if (q_len > 16) CBK_CUDA(cudaMemsetAsync(d_out_scores, 0, static_cast<size_t>(n_cand_total) * sizeof(float), stream));
  if (bf16) {
    CBK_CUDA(cudaFuncSetAttribute(maxsim_wide_stream_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    maxsim_wide_stream_kernel<__nv_bfloat16, true><<<grid, kWsThreads, smem, stream>>>(
        maps, d_pfxsum, d_doclens, n_docs, pid_base, fixed_len, skip, ss, d_Q, d_q_lens, q_len, dim, n_queries, d_cand_pids,
        d_cand_rowptr,
        n_cand_total, n_stages, idesc, d_out_scores, probe);
  } else {
    CBK_CUDA(cudaFuncSetAttribute(maxsim_wide_stream_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    maxsim_wide_stream_kernel<__half, false><<<grid, kWsThreads, smem, stream>>>(
        maps, d_pfxsum, d_doclens, n_docs, pid_base, fixed_len, skip, ss, d_Q, d_q_lens, q_len, dim, n_queries, d_cand_pids,
        d_cand_rowptr,
        n_cand_total, n_stages, idesc, d_out_scores, probe);
  }
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk
