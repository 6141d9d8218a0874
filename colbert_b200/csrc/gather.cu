// Row gather (sm_100a): `stride` consecutive store rows per pid, upcast to fp32, plus the length
// mask.  Replaces reference colbert/ranking/colbert_ranker.py:105-109 as surfaced by
// rank_forward(output_D_embedding=True) (l.131-136).  Pure copy work: 16-byte loads, 16-byte stores,
// one warp per output row, grid sized from the SM count.
#include <algorithm>

#include "cbk_common.cuh"

namespace cbk {

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const T* __restrict__ store, int64_t n_store_rows, int dim, const int64_t* __restrict__ pfxsum,
                   const int32_t* __restrict__ doclens, int64_t n_docs, const int64_t* __restrict__ pids, int64_t n,
                   int stride, float* __restrict__ out_D, uint8_t* __restrict__ out_mask) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t n_rows = n * stride;
  const int vec_per_row = dim >> 3;  // 8 16-bit elements per 16-byte load
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows;
       r += warps_total) {
    const int64_t i = r / stride;
    const int t = static_cast<int>(r - i * stride);
    const int64_t pid = pids[i];
    const bool ok = pid >= 0 && pid < n_docs;
    const int64_t src = ok ? pfxsum[pid] + t : n_store_rows;
    const int len = ok ? doclens[pid] : 0;
    float* dst = out_D + r * dim;
    for (int v = lane; v < vec_per_row; v += 32) {
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (src < n_store_rows) raw = *reinterpret_cast<const uint4*>(store + src * dim + v * 8);
      const T* e = reinterpret_cast<const T*>(&raw);
      float4 a = make_float4(to_float<T>(e[0]), to_float<T>(e[1]), to_float<T>(e[2]), to_float<T>(e[3]));
      float4 b = make_float4(to_float<T>(e[4]), to_float<T>(e[5]), to_float<T>(e[6]), to_float<T>(e[7]));
      *reinterpret_cast<float4*>(dst + v * 8) = a;
      *reinterpret_cast<float4*>(dst + v * 8 + 4) = b;
    }
    if (lane == 0) out_mask[r] = (t + 1 <= len) ? 1 : 0;
  }
}

// out[r, :] = cast(src[r, :] * mask[r])  — the multiplicative masks of BaseModel.score
// (reference colbert/modeling/BaseModel.py:41-42) fused with the cast to the MMA input type.
template <typename TIn, typename TOut, typename TMask>
__global__ void __launch_bounds__(256)
mask_cast_rows_kernel(const TIn* __restrict__ src, int64_t n_rows, int dim, const TMask* __restrict__ mask,
                      TOut* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows;
       r += warps_total) {
    const float m = mask ? static_cast<float>(mask[r]) : 1.f;
    for (int c = lane; c < dim; c += 32) out[r * dim + c] = static_cast<TOut>(static_cast<float>(src[r * dim + c]) * m);
  }
}

template <typename TIn, typename TOut>
int mask_cast_launch(const void* src, int64_t n_rows, int dim, const void* mask, int mask_dtype, void* out, int grid,
                     cudaStream_t stream) {
  const TIn* s = static_cast<const TIn*>(src);
  TOut* o = static_cast<TOut*>(out);
  switch (mask_dtype) {
    case 0: mask_cast_rows_kernel<TIn, TOut, uint8_t><<<grid, 256, 0, stream>>>(s, n_rows, dim, nullptr, o); break;
    case 1: mask_cast_rows_kernel<TIn, TOut, uint8_t><<<grid, 256, 0, stream>>>(s, n_rows, dim, static_cast<const uint8_t*>(mask), o); break;
    case 2: mask_cast_rows_kernel<TIn, TOut, int64_t><<<grid, 256, 0, stream>>>(s, n_rows, dim, static_cast<const int64_t*>(mask), o); break;
    case 3: mask_cast_rows_kernel<TIn, TOut, float><<<grid, 256, 0, stream>>>(s, n_rows, dim, static_cast<const float*>(mask), o); break;
    default: set_error("cbk_mask_cast_rows: unknown mask dtype %d", mask_dtype); return CBK_ERR_INVALID_ARG;
  }
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace

int mask_cast_dispatch(const void* d_src, int src_dtype, int64_t n_rows, int dim, const void* d_mask, int mask_dtype,
                       void* d_out, int out_dtype, cudaStream_t stream) {
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n_rows + 7) / 8, static_cast<int64_t>(sm_count()) * 8)));
#define CBK_MC(SRC, TIN)                                                                                              \
  if (src_dtype == SRC) {                                                                                             \
    if (out_dtype == CBK_F16) return mask_cast_launch<TIN, __half>(d_src, n_rows, dim, d_mask, mask_dtype, d_out, grid, stream);        \
    if (out_dtype == CBK_BF16) return mask_cast_launch<TIN, __nv_bfloat16>(d_src, n_rows, dim, d_mask, mask_dtype, d_out, grid, stream); \
    if (out_dtype == CBK_F32) return mask_cast_launch<TIN, float>(d_src, n_rows, dim, d_mask, mask_dtype, d_out, grid, stream);          \
  }
  CBK_MC(CBK_F16, __half)
  CBK_MC(CBK_BF16, __nv_bfloat16)
  CBK_MC(CBK_F32, float)
#undef CBK_MC
  set_error("cbk_mask_cast_rows: unsupported dtype pair (%d → %d)", src_dtype, out_dtype);
  return CBK_ERR_INVALID_ARG;
}

int gather_dispatch(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                    const int32_t* d_doclens, int64_t n_docs, const int64_t* d_pids, int64_t n, int stride,
                    float* d_out_D, uint8_t* d_out_mask, cudaStream_t stream) {
  const int64_t rows = n * stride;
  const int64_t blocks_wanted = (rows + 7) / 8;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(blocks_wanted, static_cast<int64_t>(sm_count()) * 8)));
  if (store_dtype == CBK_F16)
    gather_rows_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(d_store), n_store_rows, dim,
                                                         d_pfxsum, d_doclens, n_docs, d_pids, n, stride, d_out_D,
                                                         d_out_mask);
  else
    gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(d_store),
                                                                n_store_rows, dim, d_pfxsum, d_doclens, n_docs, d_pids,
                                                                n, stride, d_out_D, d_out_mask);
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk
