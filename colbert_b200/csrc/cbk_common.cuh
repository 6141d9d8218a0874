// Shared device helpers for the colbert_b200 kernels (sm_100a only): mbarrier, TMA (tensor-map
// tiled loads), ldmatrix / mma.sync wrappers, and host-side error plumbing for the C ABI.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/colbert_b200.h"

namespace cbk {

// ---------------------------------------------------------------------------------------------
// host: error plumbing (no exception crosses the C ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

#define CBK_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      ::cbk::set_error(__VA_ARGS__);              \
      return CBK_ERR_INVALID_ARG;                 \
    }                                             \
  } while (0)

#define CBK_CHECK_SUPPORTED(cond, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      ::cbk::set_error(__VA_ARGS__);              \
      return CBK_ERR_UNSUPPORTED;                 \
    }                                             \
  } while (0)

#define CBK_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::cbk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CBK_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// Encodes a 2-D tiled tensor map over a row-major [rows, dim] 16-bit matrix with a
// {box_cols, box_rows} box and 128-byte swizzle.  Returns a cbk_status.
int make_store_tensor_map(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_cols,
                          int box_rows);
// 3-D variant for dim = 128: one op loads both 64-column halves of a [box_rows, 128] tile (see api.cu)
int make_store_tensor_map_3d(CUtensorMap* out, const void* base, int64_t rows, int box_rows);

// ---------------------------------------------------------------------------------------------
// device: shared-memory address, mbarrier, TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Pins a value in a register: the compiler may not rematerialise it at every use (shared-window addresses of
// barriers otherwise get re-derived from SR_CgaCtaId inside the issue loops).
__device__ __forceinline__ uint32_t hold(uint32_t x) {
  asm volatile("" : "+r"(x));
  return x;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// One lane of a converged warp (elect.sync).  Issuing a uniform-datapath instruction (TMA load, tcgen05.mma,
// tcgen05.commit) from inside `if (elect_one())` on a warp whose control flow is otherwise uniform lets ptxas emit
// it bare, with uniform-register operands; from a `lane == 0` branch every such instruction is wrapped in an
// ELECT / R2UR / branch loop — measured with cbk_selftest_umma_rate: 154 instead of 86 cycles per 128x128x16 MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// L2 eviction-priority policies (same encodings CUTLASS uses for TMA cache hints)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// 2-D tiled TMA load: box at (x = column element, y = row) → smem, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int x, int y,
                                            uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar), "l"(policy)
      : "memory");
}

// 3-D tiled TMA load (x = column, y = row, z = 64-column half), completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, int x, int y, int z,
                                            uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3, %4}], [%5], %6;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(bar), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// device: ldmatrix / mma.sync (16-bit inputs, fp32 accumulate)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                            uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);

template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                  uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4],
                                                         uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);

template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<__half>(__half v) {
  return __half2float(v);
}
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}

// score bits → unsigned key that sorts ascending like the float.  NaN (the score of a pid outside the index, see
// cbk_maxsim_rerank) maps to 1: below -inf (0x007fffff) and above the padding key 0, so an invalid candidate sorts
// LAST instead of first; ordered_to_float(1) is again a NaN.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  if (f != f) return 1u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace cbk
