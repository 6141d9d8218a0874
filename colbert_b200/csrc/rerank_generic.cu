// Generic-dimension MaxSim (sm_100a, CUDA cores): the same contract as maxsim_rerank_kernel for any
// embedding width (the author's configuration uses dim = 768 without projection — reference
// proj_conf/dense.yaml:8 — and the reference's own known-answer vector has h = 3).  The tuned path is the
// tensor-core kernel for dim = 128; this one trades speed for generality: one CTA per query keeps the query
// (fp32, rows padded to dim+1 words against bank conflicts) in shared memory, one warp per candidate, lane =
// query row, fp32 FMAs over the document rows, which every lane reads as a broadcast.
#include <algorithm>

#include "cbk_common.cuh"

namespace cbk {

namespace {

struct StrideSet {
  int n;
  int v[CBK_MAX_STRIDES];
};

template <typename T>
__global__ void __launch_bounds__(256)
maxsim_generic_kernel(const T* __restrict__ store, int64_t n_store_rows, int dim, const int64_t* __restrict__ pfxsum,
                      const int32_t* __restrict__ doclens, int64_t n_docs, int64_t pid_base, int skip_foreign,
                      StrideSet strides, const float* __restrict__ Q, const int32_t* __restrict__ q_lens, int q_len, const int64_t* __restrict__ cand_pids,
                      const int64_t* __restrict__ rowptr, float* __restrict__ out) {
  extern __shared__ float sQ[];   // [32][dim + 1]
  const int64_t q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int pitch = dim + 1;
  const int ql = q_lens ? min(q_len, q_lens[q]) : q_len;   // rows at or past this query's own length read as zero
  for (int i = threadIdx.x; i < 32 * dim; i += blockDim.x) {
    const int r = i / dim, c = i - r * dim;
    sQ[r * pitch + c] = r < ql ? Q[(q * q_len + r) * dim + c] : 0.f;
  }
  __syncthreads();
  const int64_t beg = rowptr[q], end = rowptr[q + 1];
  const float* qrow = sQ + lane * pitch;
  for (int64_t c = beg + warp; c < end; c += n_warps) {
    const int64_t pid = cand_pids[c] - pid_base;
    if (pid < 0 || pid >= n_docs) {
      if (lane == 0) out[c] = skip_foreign ? -INFINITY : __int_as_float(0x7fc00000);
      continue;
    }
    const int64_t row0 = pfxsum[pid];
    const int len = doclens[pid];
    float best = -INFINITY;
    for (int t = 0; t < len; ++t) {
      const T* d = store + (row0 + t) * dim;
      float acc = 0.f;
      for (int k = 0; k < dim; ++k) acc = fmaf(qrow[k], to_float<T>(d[k]), acc);
      best = fmaxf(best, acc);
    }
    bool do_floor = strides.n > 0;
    for (int i = 0; i < CBK_MAX_STRIDES; ++i)
      if (i < strides.n && strides.v[i] == len) do_floor = false;
    float v = lane < q_len ? (do_floor ? fmaxf(best, 0.f) : best) : 0.f;
    if (len == 0) v = 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[c] = v;
  }
}

}  // namespace

int rerank_generic_dispatch(const void* d_store, int store_dtype, int64_t n_store_rows, int dim, const int64_t* d_pfxsum,
                            const int32_t* d_doclens, int64_t n_docs, int64_t pid_base, const int32_t* strides,
                            int n_strides, const float* d_Q, const int32_t* d_q_lens, int q_len, int64_t n_queries, const int64_t* d_cand_pids,
                            const int64_t* d_cand_rowptr, float* d_out_scores, int flags, cudaStream_t stream) {
  StrideSet ss;
  ss.n = n_strides;
  for (int i = 0; i < CBK_MAX_STRIDES; ++i) ss.v[i] = i < n_strides ? strides[i] : -1;
  const int skip = (flags & CBK_FLAG_SKIP_FOREIGN_PIDS) ? 1 : 0;
  const size_t smem = static_cast<size_t>(32) * (dim + 1) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("cbk_maxsim_rerank: dim %d too large for the generic kernel", dim);
    return CBK_ERR_UNSUPPORTED;
  }
  const unsigned int grid = static_cast<unsigned int>(n_queries);
  if (store_dtype == CBK_F16) {
    CBK_CUDA(cudaFuncSetAttribute(maxsim_generic_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    maxsim_generic_kernel<__half><<<grid, 256, smem, stream>>>(static_cast<const __half*>(d_store), n_store_rows, dim, d_pfxsum,
                                                               d_doclens, n_docs, pid_base, skip, ss, d_Q, d_q_lens, q_len, d_cand_pids,
                                                               d_cand_rowptr, d_out_scores);
  } else {
    CBK_CUDA(cudaFuncSetAttribute(maxsim_generic_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    maxsim_generic_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(d_store), n_store_rows, dim,
                                                                      d_pfxsum, d_doclens, n_docs, pid_base, skip, ss, d_Q, d_q_lens, q_len,
                                                                      d_cand_pids, d_cand_rowptr, d_out_scores);
  }
  CBK_CUDA(cudaGetLastError());
  count_launch();
  return CBK_OK;
}

}  // namespace cbk
