"""tcgen05 / TMEM / TMA building blocks (cbk_selftest_umma_gemm) against a float64 product.

Measured on B200: kind::f16 with A = fp16 and B = bf16 in ONE instruction raises cudaErrorIllegalInstruction,
so both operands must share a format (the bf16 paths split the query into hi + lo bf16 parts instead)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [16, 96, 128, 176, 256])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_umma_gemm_matches_float64(N, dt):
    from colbert_b200 import kernels
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(N)
    A = torch.randn(128, 128, generator=g).to(dt)
    B = torch.randn(N, 128, generator=g).to(dt)
    C = kernels.selftest_umma_gemm(A.to(dev), B.to(dev)).cpu().double()
    ref = A.double() @ B.double().T
    err = (C - ref).abs().max().item()
    print(f"N={N} {dt}: max abs err {err:.3e} (ref max {ref.abs().max().item():.1f})")
    assert err < 2e-3


@pytest.mark.parametrize("N", [16, 96, 256])
def test_umma_gemm_with_3d_tensor_map(N):
    """One TMA op per [rows, 128] tile: the store described as {64 columns, rows, 2 halves} with the half
    dimension having the smaller stride.  Same result as the two-op form."""
    from colbert_b200 import kernels
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(100 + N)
    A = torch.randn(128, 128, generator=g).to(torch.float16)
    B = torch.randn(N, 128, generator=g).to(torch.float16)
    C3 = kernels.selftest_umma_gemm(A.to(dev), B.to(dev), tma_3d=True).cpu().double()
    ref = A.double() @ B.double().T
    assert (C3 - ref).abs().max().item() < 2e-3


def test_umma_rate_probe_runs_and_elect_issue_is_faster():
    """cbk_selftest_umma_rate: the measurement behind the issue-path choices of the tcgen05 kernels."""
    import torch
    from colbert_b200 import kernels
    dev = torch.device("cuda", 0)
    it = 2000
    legacy = kernels.selftest_umma_rate(128, 0, it, 4, 1, dev).float().mean().item() / it
    elect = kernels.selftest_umma_rate(128, 4, it, 4, 1, dev).float().mean().item() / it
    two = kernels.selftest_umma_rate(128, 6, it, 4, 1, dev).float().mean().item() / it
    torch.cuda.synchronize()
    assert 400 < two <= elect < legacy < 3000, (legacy, elect, two)     # cycles per 128x128x128 tile; floor 512
    rd = kernels.selftest_umma_rate(16, 8, 256, 2, 1, dev)
    assert int(rd.min()) > 0
    with pytest.raises(Exception):
        kernels.selftest_umma_rate(128, 2, it, 4, 1, dev)               # two issuers without the elect path
