// Self-test of the tcgen05 building blocks (TMA SW128 tiles → tcgen05.mma → TMEM → tcgen05.ld):
//   C[128, N] = A[128, 128] · B[N, 128]^T   (16-bit inputs, fp32 accumulate), one CTA.
// Exposed as cbk_selftest_umma_gemm; the GPU tests compare it with a float64 product, so a wrong
// descriptor bit or TMEM lane mapping shows up here rather than inside the scoring kernels.
#include "umma.cuh"

namespace cbk {

namespace {

struct ProbeMaps {
  CUtensorMap a, b;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ ProbeMaps maps, int N, uint32_t idesc, float* __restrict__ C, int use_3d) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_addr = base;                 // 2 halves × 128 rows × 128 B
  const uint32_t b_addr = base + 32768;         // 2 halves × N rows × 128 B
  const uint32_t b_half = static_cast<uint32_t>(N) * 128u;
  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(N)) ncols <<= 1;

  if (tid == 0) {
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(smem_u32(&tmem_base_smem), ncols);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;

  if (tid == 0) {
    const uint32_t full = smem_u32(&bar_full);
    mbar_arrive_expect_tx(full, 32768u + 2u * b_half);
    if (use_3d) {   // one op per operand: {64 columns, rows, 2 halves} box
      tma_load_3d(a_addr, &maps.a, 0, 0, 0, full, kEvictNormal);
      tma_load_3d(b_addr, &maps.b, 0, 0, 0, full, kEvictNormal);
    } else {
      tma_load_2d(a_addr, &maps.a, 0, 0, full, kEvictNormal);
      tma_load_2d(a_addr + 16384, &maps.a, 64, 0, full, kEvictNormal);
      tma_load_2d(b_addr, &maps.b, 0, 0, full, kEvictNormal);
      tma_load_2d(b_addr + b_half, &maps.b, 64, 0, full, kEvictNormal);
    }
    mbar_wait(full, 0);
    umma::fence_after_sync();
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma::make_smem_desc_sw128(a_addr + h * 16384 + k * 32);
        const uint64_t bd = umma::make_smem_desc_sw128(b_addr + h * b_half + k * 32);
        umma::mma_f16_ss(tmem, ad, bd, idesc, (h | k) ? 1u : 0u);
      }
    umma::commit(smem_u32(&bar_mma));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar_mma), 0);
  umma::fence_after_sync();

  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    umma::tmem_ld_32x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) C[row * N + c0 + j] = __uint_as_float(r[j]);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, ncols);
}

// Issue-rate probe: every CTA multiplies resident (zero-filled) shared-memory operands `iters` times, K = 128 per
// tile (8 MMAs), rotating over n_acc TMEM accumulators; the cycles the CTA took are written out.  mode bits:
//   1  A operand from TMEM (TS) instead of shared memory (legacy issue path only)
//   4  issue through elect.sync on a converged warp (otherwise: thread 0 inside a divergent branch — the slow way,
//      kept so that the difference stays measurable)
//   2  (with 4) two issuing warps, each with half of the accumulators and half of the tiles
//   16 sixteen more warps read the accumulators back with tcgen05.ld while the MMA stream runs
__global__ void __launch_bounds__(640)
umma_rate_kernel(int N, uint32_t idesc, int mode, int iters, int n_acc, uint32_t ncols, long long* __restrict__ out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tmem_base_smem;
  __shared__ volatile int s_done;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 32768;
  const uint32_t b_half = static_cast<uint32_t>(N) * 128u;
  uint4* z = reinterpret_cast<uint4*>(smem_raw + (base - raw));
  for (uint32_t i = tid; i < (32768u + 2u * b_half) / 16u; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
    s_done = 0;
  }
  fence_proxy_async();
  if (warp == 1) {
    umma::tmem_alloc(smem_u32(&tmem_base_smem), ncols);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;
  const int n_issuers = (mode & 6) == 6 ? 2 : 1;

  if (warp >= 4) {
    // the interference an epilogue causes: tcgen05.ld + max tree for as long as the MMA stream runs
    float acc = 0.f;
    const uint32_t rd = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    int i = 0;
    while (!s_done) {
      uint32_t v[32];
      umma::tmem_ld_32x32(rd + ((i++ * 32) & (ncols - 1)), v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 2) acc = fmaxf(acc, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
    }
    if (acc == 12345.678f) out_cycles[blockIdx.x] = 0;   // keeps the loads alive
  } else if ((mode & 4) && warp < n_issuers) {
    // converged-warp issue: the whole warp walks the loop, one elected lane issues
    const int my_acc = n_acc / n_issuers, my_iters = iters / n_issuers;
    const long long t0 = clock64();
    for (int i = 0; i < my_iters; ++i) {
      const int slot = warp * my_acc + i % my_acc;
      if (i >= my_acc) mbar_wait(smem_u32(&bars[slot]), ((i / my_acc) - 1) & 1);
      umma::fence_after_sync();
      const uint32_t d = tmem + static_cast<uint32_t>(slot * N);
      if (elect_one()) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma::mma_f16_ss(d, umma::make_smem_desc_sw128(a_addr + h * 16384 + k * 32),
                             umma::make_smem_desc_sw128(b_addr + h * b_half + k * 32), idesc, (h | k) ? 1u : 0u);
        umma::commit(smem_u32(&bars[slot]));
      }
      __syncwarp();
    }
    for (int j = 0; j < my_acc && j < my_iters; ++j) {   // last use of each of my accumulators
      const int uses = (my_iters - j + my_acc - 1) / my_acc;
      mbar_wait(smem_u32(&bars[warp * my_acc + j]), (uses - 1) & 1);
    }
    if (tid == 0) {
      out_cycles[blockIdx.x] = clock64() - t0;
      s_done = 1;
    }
  } else if (!(mode & 4) && tid == 0) {
    // legacy issue path: one thread inside a divergent branch (ptxas wraps every MMA in an ELECT / R2UR loop)
    const uint32_t a_tmem = tmem + ncols - 64;     // TS mode: 128 lanes × 64 columns hold A[128, 128] 16-bit
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int slot = i % n_acc;
      if (i >= n_acc) mbar_wait(smem_u32(&bars[slot]), ((i / n_acc) - 1) & 1);
      umma::fence_after_sync();
      const uint32_t d = tmem + static_cast<uint32_t>(slot * N);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t bd = umma::make_smem_desc_sw128(b_addr + h * b_half + k * 32);
          if (mode & 1) {
            umma::mma_f16_ts(d, a_tmem + (h * 4 + k) * 8, bd, idesc, (h | k) ? 1u : 0u);
          } else {
            umma::mma_f16_ss(d, umma::make_smem_desc_sw128(a_addr + h * 16384 + k * 32), bd, idesc, (h | k) ? 1u : 0u);
          }
        }
      umma::commit(smem_u32(&bars[slot]));
    }
    for (int slot = 0; slot < n_acc && slot < iters; ++slot) {
      const int uses = (iters - slot + n_acc - 1) / n_acc;
      mbar_wait(smem_u32(&bars[slot]), (uses - 1) & 1);
    }
    out_cycles[blockIdx.x] = clock64() - t0;
    s_done = 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, ncols);
}

// TMEM read-rate probe: `blockDim.x / 32` warps (warp w reads lane quadrant w % 4) each issue `iters` tcgen05.ld
// of 32 lanes x 32 columns (4 KB), `depth` loads between waits; cycles of warp 0 are written out.
__global__ void __launch_bounds__(512)
tmem_ld_rate_kernel(int iters, int depth, long long* __restrict__ out_cycles) {
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    umma::tmem_alloc(smem_u32(&tmem_base_smem), 512);
    umma::tmem_relinquish();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((warp >> 2) * 128);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += depth) {
    uint32_t v[2][32];
    umma::tmem_ld_32x32(base + (i & 3) * 32, v[0]);
    if (depth > 1) umma::tmem_ld_32x32(base + ((i + 1) & 3) * 32, v[1]);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 2) acc = fmaxf(acc, fmaxf(__uint_as_float(v[0][j]), __uint_as_float(v[0][j + 1])));
    if (depth > 1) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) acc = fmaxf(acc, fmaxf(__uint_as_float(v[1][j]), __uint_as_float(v[1][j + 1])));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) out_cycles[blockIdx.x] = 0;   // keep the loads alive
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

int umma_rate_dispatch(int N, int mode, int iters, int n_acc, int ctas_per_sm, long long* d_cycles, cudaStream_t stream) {
  if (mode == 8) {   // TMEM read rate: N = warps per CTA (4, 8, 12, 16), n_acc = loads between waits (1 or 2)
    CBK_CHECK_ARG((N == 4 || N == 8 || N == 12 || N == 16) && (n_acc == 1 || n_acc == 2) && iters >= 2 && d_cycles,
                  "cbk_selftest_umma_rate: bad TMEM-read arguments");
    tmem_ld_rate_kernel<<<sm_count(), N * 32, 0, stream>>>(iters, n_acc, d_cycles);
    CBK_CUDA(cudaGetLastError());
    count_launch();
    return CBK_OK;
  }
  CBK_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0 && iters >= 2 && n_acc >= 1 && n_acc <= 8 && mode >= 0 && mode <= 23 &&
                    ctas_per_sm >= 1 && ctas_per_sm <= 4 && d_cycles,
                "cbk_selftest_umma_rate: bad arguments");
  CBK_CHECK_ARG((mode & 6) != 2 && ((mode & 6) != 6 || (n_acc % 2 == 0 && iters % 2 == 0)) && (!(mode & 1) || !(mode & 4)),
                "cbk_selftest_umma_rate: mode %d: two issuers need the elect path and even n_acc / iters; TS needs the legacy path",
                mode);
  const int need = n_acc * N + ((mode & 1) ? 64 : 0);
  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(need)) ncols <<= 1;
  CBK_CHECK_ARG(ncols * ctas_per_sm <= 512, "cbk_selftest_umma_rate: %d TMEM columns x %d CTAs exceed 512", ncols, ctas_per_sm);
  const uint32_t idesc = umma::make_idesc(128, static_cast<uint32_t>(N), umma::kFmtF16, umma::kFmtF16);
  const size_t smem = 32768 + static_cast<size_t>(N) * 256 + 1024;
  CBK_CHECK_ARG(smem * ctas_per_sm <= 220 * 1024, "cbk_selftest_umma_rate: shared memory");
  CBK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  umma_rate_kernel<<<sm_count() * ctas_per_sm, (mode & 16) ? 640 : 128, smem, stream>>>(N, idesc, mode & 7, iters, n_acc, ncols, d_cycles);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

int umma_probe_dispatch(const void* d_A, const void* d_B, int N, int a_bf16, int b_bf16, float* d_C, cudaStream_t stream) {
  ProbeMaps maps;
  const int use_3d = (a_bf16 | b_bf16) & 2 ? 1 : 0;   // bit 1 of either format flag: load each operand with one 3-D TMA op
  a_bf16 &= 1;
  b_bf16 &= 1;
  int rc = use_3d ? make_store_tensor_map_3d(&maps.a, d_A, 128, 128) : make_store_tensor_map(&maps.a, d_A, 128, 128, 64, 128);
  if (rc != CBK_OK) return rc;
  rc = use_3d ? make_store_tensor_map_3d(&maps.b, d_B, N, N) : make_store_tensor_map(&maps.b, d_B, N, 128, 64, N);
  if (rc != CBK_OK) return rc;
  const uint32_t idesc = umma::make_idesc(128, static_cast<uint32_t>(N), a_bf16 ? umma::kFmtBF16 : umma::kFmtF16,
                                          b_bf16 ? umma::kFmtBF16 : umma::kFmtF16);
  const size_t smem = 32768 + static_cast<size_t>(N) * 256 + 1024;
  CBK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  umma_probe_kernel<<<1, 128, smem, stream>>>(maps, N, idesc, d_C, use_3d);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk

// ---- entry points of libcolbert_b200_probe.so (include/colbert_b200_probe.h) ---------------------------------------
static int probe_check_device() {
  int dev = 0, major = 0;
  CBK_CUDA(cudaGetDevice(&dev));
  CBK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CBK_CHECK_SUPPORTED(major == 10, "device %d has compute capability major %d; the probes are sm_100a only", dev, major);
  return CBK_OK;
}

extern "C" {

int cbk_selftest_umma_rate(int N, int mode, int iters, int n_acc, int ctas_per_sm, int64_t* d_cycles, void* stream) {
  int rc = probe_check_device();
  if (rc != CBK_OK) return rc;
  return cbk::umma_rate_dispatch(N, mode, iters, n_acc, ctas_per_sm, reinterpret_cast<long long*>(d_cycles),
                                 static_cast<cudaStream_t>(stream));
}

int cbk_selftest_umma_gemm(const void* d_A, const void* d_B, int N, int a_bf16, int b_bf16, float* d_C, void* stream) {
  CBK_CHECK_ARG(d_A && d_B && d_C, "cbk_selftest_umma_gemm: null pointer argument");
  CBK_CHECK_SUPPORTED(N >= 16 && N <= 256 && N % 16 == 0, "cbk_selftest_umma_gemm: N %d must be a multiple of 16 in [16, 256]", N);
  int rc = probe_check_device();
  if (rc != CBK_OK) return rc;
  return cbk::umma_probe_dispatch(d_A, d_B, N, a_bf16, b_bf16, d_C, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
