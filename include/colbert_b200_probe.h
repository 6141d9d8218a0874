/*
 * colbert_b200_probe.h — self-test and issue-rate probes of the tcgen05 / TMEM / TMA building blocks the tensor-core
 * kernels of libcolbert_b200.so are made of.  NOT part of the scoring path and not part of the product library: they live
 * in libcolbert_b200_probe.so (which links against libcolbert_b200.so for error reporting and launch accounting) and are
 * used by tests/test_gpu_umma.py and benchmarks/umma_rate.py.  Conventions as in colbert_b200.h.
 */
#ifndef COLBERT_B200_PROBE_H
#define COLBERT_B200_PROBE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Self-test of the tcgen05 / TMEM / TMA building blocks the query-batched kernels are made of:
 *     C[128, N] = A[128, 128] · B[N, 128]^T      16-bit inputs (a_bf16 / b_bf16: 0 = fp16, 1 = bf16), fp32 out
 * N a multiple of 16 in [16, 256]; one CTA.  Bit 1 of a_bf16 selects the 3-D tensor-map variant (one TMA op
 * per operand instead of one per 64-column half).  Not part of the scoring path.
 * ------------------------------------------------------------------------------------------------ */
int cbk_selftest_umma_gemm(const void* d_A, const void* d_B, int N, int a_bf16, int b_bf16, float* d_C, void* stream);

/* Issue-rate probe for the same building block: ctas_per_sm CTAs per SM each run `iters` tiles of
 * [128, N] += A[128, 128] · B[N, 128]^T over resident shared-memory operands, rotating over n_acc TMEM
 * accumulators; d_cycles [n_SMs * ctas_per_sm] int64 receives the clock cycles each CTA took.  mode bits:
 * 1 = A operand from TMEM, 4 = issue through elect.sync on a converged warp (otherwise thread 0 in a divergent
 * branch), 2 = (with 4) two issuing warps, 16 = sixteen more warps read the accumulators back meanwhile.
 * mode 8: TMEM read rate instead — N = warps per CTA (4/8/12/16), n_acc = loads between waits (1/2).
 * Used to place the tensor-bound kernels against what the MMA shape itself can sustain
 * (benchmarks/umma_rate.py).  Not part of the scoring path. */
int cbk_selftest_umma_rate(int N, int mode, int iters, int n_acc, int ctas_per_sm, int64_t* d_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COLBERT_B200_PROBE_H */
