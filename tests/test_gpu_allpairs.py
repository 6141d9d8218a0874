"""All-pairs MaxSim with a backward pass (cbk_score_allpairs_fwd / _bwd, csrc/score_allpairs.cu) — BaseModel.score as the
reference trains with it (BaseModel.py:39-46 under autograd, colbert_model.py:87-95).

Parity ladder: (1) goldens produced by the reference's own ops under torch autograd (tests/golden/score_grad_cases.npz);
(2) the numpy oracle at shapes it finishes in seconds; (3) at the training shape a plain torch fp32 einsum on the GPU (this
is a floating-point kernel: tolerance 1e-3 relative, the north-star's bound for 16-bit operands)."""
import os

import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as O
from parity_utils import SCORE_RTOL

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-3   # relative to the largest gradient entry of the tensor (operands are rounded to 16 bits once)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def _score_with_grads(Q, D, qmask, dmask, W, dev, store_dtype=torch.float16):
    from colbert_b200.modeling.BaseModel import BaseModel
    Qt = torch.from_numpy(np.asarray(Q, dtype=np.float32)).to(dev).requires_grad_(True)
    Dt = torch.from_numpy(np.asarray(D, dtype=np.float32)).to(dev).requires_grad_(True)
    s = BaseModel.score(Qt, Dt, torch.from_numpy(qmask).to(dev), torch.from_numpy(dmask).to(dev), store_dtype=store_dtype)
    (s * torch.from_numpy(W).to(dev)).sum().backward()
    return s.detach().cpu().numpy(), Qt.grad.cpu().numpy(), Dt.grad.cpu().numpy()


def _rel(a, b):
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


@pytest.mark.parametrize("name", ["g_small", "g_mid", "g_views", "g_wide", "g_tiles"])
def test_score_and_gradients_match_the_reference_goldens(golden_dir, dev, name):
    g = np.load(os.path.join(golden_dir, "score_grad_cases.npz"))
    s, dQ, dD = _score_with_grads(g[f"{name}_Q"], g[f"{name}_D"], g[f"{name}_qmask"], g[f"{name}_dmask"], g[f"{name}_W"], dev)
    ref = g[f"{name}_score"]
    assert np.abs(s - ref).max() <= SCORE_RTOL * max(1.0, np.abs(ref).max())
    # the goldens' inputs are fp16 values, so the 16-bit operands are exact and only the summation order differs
    assert _rel(dQ, g[f"{name}_dQ"]) <= 1e-5, _rel(dQ, g[f"{name}_dQ"])
    assert _rel(dD, g[f"{name}_dD"]) <= 1e-5, _rel(dD, g[f"{name}_dD"])


def _random_case(rng, nq, m, nd, n, h, full_masks=False):
    Q = rng.standard_normal((nq, m, h), dtype=np.float32)
    D = rng.standard_normal((nd, n, h), dtype=np.float32)
    Q /= np.linalg.norm(Q, axis=-1, keepdims=True)
    D /= np.linalg.norm(D, axis=-1, keepdims=True)
    qlen = np.full(nq, m) if full_masks else rng.integers(1, m + 1, size=nq)
    dlen = np.full(nd, n) if full_masks else rng.integers(1, n + 1, size=nd)
    qmask = (np.arange(m)[None, :] < qlen[:, None]).astype(np.int64)
    dmask = (np.arange(n)[None, :] < dlen[:, None]).astype(np.int64)
    W = rng.standard_normal((nq, nd), dtype=np.float32)
    return Q, D, qmask, dmask, W


@pytest.mark.parametrize("shape", [(1, 1, 1, 1, 64), (7, 32, 5, 384, 128), (9, 20, 33, 180, 192), (13, 32, 40, 16, 768),
                                   (4, 32, 3, 700, 64), (5, 3, 257, 1, 64), (6, 32, 2, 256, 1024)])
def test_against_the_oracle(dev, shape):
    """odd query counts (partial query block), m < 32, documents that straddle 256-row tiles, n = 1, widths 64 … 1024"""
    rng = np.random.default_rng(sum(shape))
    Q, D, qmask, dmask, W = _random_case(rng, *shape)
    s, dQ, dD = _score_with_grads(Q, D, qmask, dmask, W, dev)
    Q16, D16 = Q.astype(np.float16).astype(np.float32), D.astype(np.float16).astype(np.float32)
    rs, rdQ, rdD, _ = O.score_allpairs_grad(Q16, D16, qmask, dmask, W)
    assert np.abs(s - rs).max() <= 1e-4 * max(1.0, np.abs(rs).max())          # same 16-bit operands: accumulation order only
    ex = O.score_allpairs(Q, D, qmask, dmask)                                     # and the north-star bound vs exact fp32 inputs
    assert np.abs(s - ex).max() <= SCORE_RTOL * max(1.0, np.abs(ex).max())
    assert _rel(dQ, rdQ) <= 1e-4, _rel(dQ, rdQ)
    assert _rel(dD, rdD) <= 1e-4, _rel(dD, rdD)


def test_all_rows_masked_and_empty_masks(dev):
    """a fully masked document scores exactly 0 against everything (multiplicative mask, BaseModel.py:41) and gets no gradient"""
    rng = np.random.default_rng(3)
    Q, D, qmask, dmask, W = _random_case(rng, 5, 32, 6, 50, 128)
    dmask[2] = 0
    qmask[1] = 0
    s, dQ, dD = _score_with_grads(Q, D, qmask, dmask, W, dev)
    assert np.all(s[:, 2] == 0.0) and np.all(s[1] == 0.0)
    assert np.all(dD[2] == 0.0) and np.all(dQ[1] == 0.0)
    assert np.all(dD[dmask == 0] == 0.0) and np.all(dQ[qmask == 0] == 0.0)


def test_known_answer_vector_padded_to_64(dev):
    """BaseModel.test_score (BaseModel.py:70-75) with its 3-wide vectors zero-padded to the tensor-core width → [[21, 41]]"""
    from colbert_b200.modeling.BaseModel import BaseModel
    Q = torch.zeros(1, 2, 64)
    D = torch.zeros(2, 2, 64)
    Q[0, :, :3] = torch.tensor([[1., 5, 4], [2, 8, 1]])
    D[:, :, :3] = torch.tensor([[[0., 0, 0], [1, 1, 1]], [[3, 2, 1], [1, 1, 3]]])
    s = BaseModel.score(Q.to(dev), D.to(dev), torch.ones(1, 2).to(dev), torch.ones(2, 2).to(dev))
    assert s.cpu().tolist() == [[21.0, 41.0]]


def test_gradients_are_bit_reproducible(dev):
    rng = np.random.default_rng(11)
    case = _random_case(rng, 24, 32, 30, 200, 128)
    a = _score_with_grads(*case, dev)
    b = _score_with_grads(*case, dev)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_training_shape_against_torch_fp32(dev):
    """a slice of the author's training batch (h = 768, m = 32, n = 384; colbert_model.py:87-95) against the reference's
    own op sequence in fp32 on the GPU.  An fp32 matmul and the tensor cores add in different orders, so a near-tie may
    resolve to another row: the kernel's arg-max has to BE a maximum (within 1e-5), and the gradients have to be what
    autograd derives for that arg-max (gather for dQ, index_add for dD)."""
    from colbert_b200 import kernels
    from colbert_b200.modeling.BaseModel import BaseModel
    torch.manual_seed(5)
    nq, m, nd, n, h = 34, 32, 68, 384, 768
    Q = torch.nn.functional.normalize(torch.randn(nq, m, h, device=dev), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(nd, n, h, device=dev), dim=-1)
    qmask = (torch.arange(m, device=dev)[None, :] < torch.randint(8, m + 1, (nq, 1), device=dev)).long()
    dmask = (torch.arange(n, device=dev)[None, :] < torch.randint(20, n + 1, (nd, 1), device=dev)).long()
    W = torch.randn(nq, nd, device=dev)
    Qa, Da = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
    s = BaseModel.score(Qa, Da, qmask, dmask)
    (s * W).sum().backward()
    Qp = kernels.mask_cast_rows(Q.reshape(-1, h), qmask.reshape(-1), torch.float16).reshape(nq, m, h)
    Dp = kernels.mask_cast_rows(D.reshape(-1, h), dmask.reshape(-1), torch.float16).reshape(nd, n, h)
    s2, arg = kernels.score_allpairs_fwd(Qp, Dp)
    assert torch.equal(s2, s.detach())
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        Qm, Dm = Qp.float(), Dp.float()
        sim = torch.einsum("qmh,dnh->qdmn", Qm, Dm)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    mx = sim.max(-1)[0]
    ref = mx.sum(-1)
    assert float((s.detach() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    at = sim.gather(-1, arg.long().unsqueeze(-1)).squeeze(-1)
    assert float((mx - at).max()) <= 1e-5
    rows = ((torch.arange(nd, device=dev) * n)[None, :, None] + arg.long()).reshape(-1)
    dQ_ref = (Dm.reshape(nd * n, h)[rows].view(nq, nd, m, h) * W[:, :, None, None]).sum(1) * qmask[..., None]
    dD_ref = torch.zeros(nd * n, h, device=dev).index_add_(0, rows, (W[:, :, None, None] * Qm[:, None]).reshape(-1, h))
    dD_ref = dD_ref.view(nd, n, h) * dmask[..., None]
    for got, want in ((Qa.grad, dQ_ref), (Da.grad, dD_ref)):
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max())


def test_forward_only_without_grad_and_bf16_operands(dev):
    from colbert_b200.modeling.BaseModel import BaseModel
    rng = np.random.default_rng(21)
    Q, D, qmask, dmask, _ = _random_case(rng, 6, 32, 10, 120, 256)
    ex = O.score_allpairs(Q, D, qmask, dmask)
    with torch.no_grad():
        s = BaseModel.score(torch.from_numpy(Q).to(dev), torch.from_numpy(D).to(dev), torch.from_numpy(qmask).to(dev),
                            torch.from_numpy(dmask).to(dev), store_dtype=torch.bfloat16)
    # bf16 operands keep 8 significant bits: the bound is the one the bf16-native rerank path documents (≈ 1e-2 relative)
    assert np.abs(s.cpu().numpy() - ex).max() <= 1e-2 * max(1.0, np.abs(ex).max())


def test_unsupported_shapes(dev):
    from colbert_b200 import kernels
    from colbert_b200._lib import CbkError
    Qp = torch.zeros(2, 33, 64, dtype=torch.float16, device=dev)
    Dp = torch.zeros(2, 8, 64, dtype=torch.float16, device=dev)
    with pytest.raises(CbkError):
        kernels.score_allpairs_fwd(Qp, Dp)
    Qp = torch.zeros(2, 8, 96, dtype=torch.float16, device=dev)
    Dp = torch.zeros(2, 8, 96, dtype=torch.float16, device=dev)
    with pytest.raises(CbkError):
        kernels.score_allpairs_fwd(Qp, Dp)
