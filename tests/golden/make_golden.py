#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) on seeded
synthetic indexes.  Runs only in the authoring container (the GPU box has no /root/reference);
the resulting small fixtures are committed and are what the tests read.

Reference entry points exercised (no edits to reference files; four import-time shims):
  * colbert.modeling.BaseModel.BaseModel.score            (BaseModel.py:39-46)
  * colbert.ranking.colbert_ranker.ColbertRanker          (colbert_ranker.py:15-137)
      – __init__/_load_parts/init_ranker on an on-disk index written in the reference layout
      – rank_forward(Q[1,dim,q_len], pids, depth, output_D_embedding)

Shims: (1) stub `faiss` module (top-level import only used by ColbertIndex); (2) `ujson` → stdlib json;
(3) DEVICE='cpu' and Tensor.cuda → identity; (4) torch.zeros without pin_memory/device='cuda'.

Usage:  python tests/golden/make_golden.py        (writes next to this file)
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

sys.modules["faiss"] = types.ModuleType("faiss")      # shim 1
sys.modules["ujson"] = json                           # shim 2
import torch  # noqa: E402

_orig_zeros = torch.zeros


def _zeros(*a, **k):                                  # shim 4
    k.pop("pin_memory", None)
    if str(k.get("device", "cpu")).startswith("cuda"):
        k["device"] = "cpu"
    return _orig_zeros(*a, **k)


torch.zeros = _zeros
torch.Tensor.cuda = lambda self, *a, **k: self        # shim 3b

import colbert.ranking.colbert_ranker as cr          # noqa: E402
from colbert.modeling.BaseModel import BaseModel      # noqa: E402

cr.DEVICE = "cpu"                                     # shim 3a

from colbert_b200 import synthetic                   # noqa: E402
from golden_cases import CASES, build_case            # noqa: E402


def run_reference_case(case):
    index, queries, cand_lists = build_case(case)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_index(index, d)
        ranker = cr.ColbertRanker(d, model=BaseModel, dim=index.dim)
        out["strides"] = np.asarray(ranker.strides, dtype=np.int64)
        out["pfxsum_tail"] = ranker.doclens_pfxsum[-4:].numpy().astype(np.int64)
        out["store_rows"] = np.asarray([ranker.tensor.shape[0]], dtype=np.int64)
        for qi, (Q, pids) in enumerate(zip(queries, cand_lists)):
            Qt = torch.from_numpy(Q).unsqueeze(0).permute(0, 2, 1)     # [1, dim, q_len] as faiss_indexers.py:232-233
            for depth_name, depth in case["depths"]:
                p, s = ranker.rank_forward(Qt, [int(x) for x in pids], depth=depth)
                out[f"q{qi}_{depth_name}_pids"] = np.asarray(p, dtype=np.int64)
                out[f"q{qi}_{depth_name}_scores"] = np.asarray(s, dtype=np.float32)
            if case.get("output_D"):
                p, D, M = ranker.rank_forward(Qt, [int(x) for x in pids], depth=case["output_D"],
                                              output_D_embedding=True)
                out[f"q{qi}_D_pids"] = np.asarray(p, dtype=np.int64)
                out[f"q{qi}_D_rows"] = D.numpy().astype(np.float16)   # exact: rows are fp16 values upcast
                out[f"q{qi}_D_mask"] = M.numpy()
    return out


def run_score_cases():
    out = {}
    # the one known-answer vector in the reference tree: BaseModel.test_score (BaseModel.py:70-75)
    Q = torch.tensor([[[1, 5, 4], [2, 8, 1]]]).float()
    D = torch.tensor([[[0, 0, 0], [1, 1, 1]], [[3, 2, 1], [1, 1, 3]]]).float()
    qm, dm = torch.ones(Q.size()[:2]), torch.ones(D.size()[:2])
    out["kat_Q"], out["kat_D"] = Q.numpy(), D.numpy()
    out["kat_score"] = BaseModel.score(Q, D, qm, dm).numpy()
    # seeded all-pairs cases with 0/1 masks (training/eval shape, colbert_model.py:90)
    rng = np.random.default_rng(77)
    for name, (nq, m, nd, n, h) in {"ap_small": (3, 5, 4, 7, 16), "ap_mid": (6, 32, 12, 40, 128),
                                    "ap_views": (4, 8, 9, 8, 128)}.items():
        Qn = rng.standard_normal((nq, m, h), dtype=np.float32)
        Dn = rng.standard_normal((nd, n, h), dtype=np.float32)
        Qn /= np.linalg.norm(Qn, axis=-1, keepdims=True)
        Dn /= np.linalg.norm(Dn, axis=-1, keepdims=True)
        qlen = rng.integers(1, m + 1, size=nq)
        dlen = rng.integers(1, n + 1, size=nd)
        qmask = (np.arange(m)[None, :] < qlen[:, None]).astype(np.int64)
        dmask = (np.arange(n)[None, :] < dlen[:, None]).astype(np.int64)
        s = BaseModel.score(torch.from_numpy(Qn), torch.from_numpy(Dn),
                            torch.from_numpy(qmask), torch.from_numpy(dmask)).numpy()
        out[f"{name}_Q"], out[f"{name}_D"] = Qn.astype(np.float16), Dn.astype(np.float16)
        # scores are regenerated from the fp16-rounded inputs so the fixture stays small and exact
        s = BaseModel.score(torch.from_numpy(out[f"{name}_Q"].astype(np.float32)),
                            torch.from_numpy(out[f"{name}_D"].astype(np.float32)),
                            torch.from_numpy(qmask), torch.from_numpy(dmask)).numpy()
        out[f"{name}_qmask"], out[f"{name}_dmask"], out[f"{name}_score"] = qmask, dmask, s
    return out


def run_score_grad_cases():
    """BaseModel.score under autograd, as the reference trains with it (colbert_model.py:87-95): scores and the gradients
    of Σ scores * W with respect to Q and D, from the reference's own ops on the CPU in fp32."""
    out = {}
    rng = np.random.default_rng(4242)
    for name, (nq, m, nd, n, h) in {"g_small": (3, 5, 4, 7, 64), "g_mid": (6, 32, 12, 40, 128),
                                    "g_views": (4, 16, 9, 16, 128), "g_wide": (2, 8, 3, 40, 768),
                                    "g_tiles": (5, 32, 3, 300, 64)}.items():
        Qn = rng.standard_normal((nq, m, h), dtype=np.float32)
        Dn = rng.standard_normal((nd, n, h), dtype=np.float32)
        Qn /= np.linalg.norm(Qn, axis=-1, keepdims=True)
        Dn /= np.linalg.norm(Dn, axis=-1, keepdims=True)
        Q16, D16 = Qn.astype(np.float16), Dn.astype(np.float16)      # inputs are stored (and differentiated) at fp16 values
        qlen = rng.integers(1, m + 1, size=nq)
        dlen = rng.integers(1, n + 1, size=nd)
        qmask = (np.arange(m)[None, :] < qlen[:, None]).astype(np.int64)
        dmask = (np.arange(n)[None, :] < dlen[:, None]).astype(np.int64)
        W = rng.standard_normal((nq, nd), dtype=np.float32)
        Qt = torch.from_numpy(Q16.astype(np.float32)).requires_grad_(True)
        Dt = torch.from_numpy(D16.astype(np.float32)).requires_grad_(True)
        s = BaseModel.score(Qt, Dt, torch.from_numpy(qmask), torch.from_numpy(dmask))
        (s * torch.from_numpy(W)).sum().backward()
        out[f"{name}_Q"], out[f"{name}_D"], out[f"{name}_qmask"], out[f"{name}_dmask"] = Q16, D16, qmask, dmask
        out[f"{name}_W"], out[f"{name}_score"] = W, s.detach().numpy()
        out[f"{name}_dQ"], out[f"{name}_dD"] = Qt.grad.numpy(), Dt.grad.numpy()
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "score_cases.npz"), **run_score_cases())
    print("wrote score_cases.npz")
    np.savez_compressed(os.path.join(HERE, "score_grad_cases.npz"), **run_score_grad_cases())
    print("wrote score_grad_cases.npz")
    for case in CASES:
        res = run_reference_case(case)
        path = os.path.join(HERE, f"rank_{case['name']}.npz")
        np.savez_compressed(path, **res)
        print("wrote", path, os.path.getsize(path), "bytes; strides", res["strides"].tolist())


if __name__ == "__main__":
    main()
