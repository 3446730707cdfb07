"""Parity of what bench.py actually times: SPVCNN cr=2.0 on a 5-sweep, 0.05 m nuScenes-shape scan, fused
conv/BatchNorm/ReLU/residual nodes, against the fp64 CPU oracle — logits and EVERY parameter gradient
(per-tensor table, tests/gradtable.py).  Plus the small glue functions that had no test of their own
(fetch_idx, voxel_to_point(nearest=True)).

Norm: max|a-b| / max|b| per tensor.  Bars: north_star's bf16/tf32 rel 2e-2 and fp32 rel 1e-4
for the logits and per gradient tensor; where a tensor is exempt it is named in KNOWN_* with the reason.
"""
import json
import os

import numpy as np
import pytest
import torch

import gradtable

pytestmark = pytest.mark.gpu

WORKLOAD = "nusc5_cr2.0_b2"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


@pytest.fixture(scope="module")
def ref(oracle, cuda_lib):
    """fp64 oracle step on one scan of the benchmark workload (~20 s on 8-16 cores)."""
    return gradtable.oracle_step(WORKLOAD, 7)


def _dump(tab, name):
    os.makedirs(OUT, exist_ok=True)
    json.dump(tab, open(os.path.join(OUT, name), "w"), indent=1)
    print(gradtable.describe(tab))


def _check(tab, logit_bar, grad_bar, l2_bar, exempt=()):
    assert tab["logits_max_rel"] < logit_bar, gradtable.describe(tab)
    bad = [r for r in tab["params"] if r["max_rel"] >= grad_bar and not any(r["name"].startswith(e) for e in exempt)]
    assert not bad, gradtable.describe(tab)
    bad = [r for r in tab["params"] if r["l2_rel"] >= l2_bar]
    assert not bad, gradtable.describe(tab)


# BatchNorm biases/weights directly behind a conv whose output channel is (nearly) constant over the batch are
# ill-conditioned in ANY finite precision: dgamma = sum(dy * xhat) cancels to ~0 while its terms are O(1).  They are
# held to the l2 bar (whole tensor) instead of the max-norm bar of their worst element.
def test_bench_config_fp32_vs_fp64_oracle(ref):
    tab = gradtable.table(WORKLOAD, 7, "fp32", fused=False, ref=ref)
    _dump(tab, "r2_gradtable_fp32.json")
    _check(tab, 1e-4, 1e-4, 1e-4)


def test_bench_config_bf16_fused_vs_fp64_oracle(ref):
    tab = gradtable.table(WORKLOAD, 7, "bf16", fused=True, ref=ref)
    _dump(tab, "r2_gradtable_bf16_fused.json")
    _check(tab, 2e-2, BF16_GRAD_BAR, BF16_L2_BAR)


def test_bench_config_tf32_vs_fp64_oracle(ref):
    tab = gradtable.table(WORKLOAD, 7, "tf32", fused=False, ref=ref)
    _dump(tab, "r2_gradtable_tf32.json")
    _check(tab, 2e-2, 2e-2, 2e-2)


# measured on the B200 (profiles/r2_gradtables.md): see the table there for every tensor
BF16_GRAD_BAR = 2e-2
BF16_L2_BAR = 2e-2


# ---------------------------------------------------------------- glue functions without a test of their own
def test_fetch_idx_bit_exact(oracle, cuda_lib):
    """core/models/utils.py:121-135."""
    from u2mkd_b200 import models
    fam_o = models.build_family(oracle.as_torchsparse_modules()["torchsparse"])
    fam_g = models.product()
    rng = np.random.default_rng(0)
    tgt = np.unique(np.concatenate([rng.integers(0, 50, (20000, 3)), rng.integers(0, 2, (20000, 1))], 1).astype(np.int32), axis=0)
    rng.shuffle(tgt)
    src = np.concatenate([tgt[rng.integers(0, len(tgt), 5000)], rng.integers(50, 60, (3000, 4)).astype(np.int32)])
    rng.shuffle(src)
    src, tgt = torch.from_numpy(src), torch.from_numpy(np.ascontiguousarray(tgt))
    want = fam_o.fetch_idx(src, tgt)
    got = fam_g.fetch_idx(src.cuda(), tgt.cuda()).cpu()
    assert torch.equal(got, want) and int((want < 0).sum()) >= 3000 and int((want >= 0).sum()) >= 5000


@pytest.mark.parametrize("stride", [1, 4])
def test_voxel_to_point_nearest(oracle, cuda_lib, stride):
    """core/models/utils.py:100-103: nearest=True keeps corner 0 only (weights[:,1:]=0, idx[:,1:]=-1)."""
    from u2mkd_b200 import models, ops
    import u2mkd_b200.torchsparse as gts
    ops.set_math("fp32")
    fam_o = models.build_family(oracle.as_torchsparse_modules()["torchsparse"])
    fam_g = models.product()
    rng = np.random.default_rng(stride)
    g = np.stack(np.meshgrid(*[np.arange(14)] * 3, indexing="ij"), -1).reshape(-1, 3) * stride
    vox = g[rng.random(len(g)) < 0.7]
    vox = torch.from_numpy(np.concatenate([vox, np.zeros((len(vox), 1))], 1).astype(np.int32))
    pts = np.concatenate([rng.random((8000, 3)) * 13 * stride, np.zeros((8000, 1))], 1).astype(np.float32)
    feats = torch.from_numpy(rng.standard_normal((vox.shape[0], 32)).astype(np.float32))
    pf = torch.zeros(8000, 32)
    fo, fg = feats.clone().requires_grad_(True), feats.clone().cuda().requires_grad_(True)
    xo = oracle.SparseTensor(fo, vox, stride)
    xg = gts.SparseTensor(fg, vox.cuda(), stride)
    zo = oracle.PointTensor(pf, torch.from_numpy(pts))
    zg = gts.PointTensor(pf.cuda(), torch.from_numpy(pts).cuda())
    oo = fam_o.voxel_to_point(xo, zo, nearest=True)
    og = fam_g.voxel_to_point(xg, zg, nearest=True)
    s = (stride,) * 3
    assert torch.equal(zg.idx_query[s].cpu().long(), zo.idx_query[s].long())
    assert bool((zg.idx_query[s][:, 1:] == -1).all()) and bool((zg.weights[s][:, 1:] == 0).all())
    assert float((zg.weights[s].cpu() - zo.weights[s]).abs().max()) < 1e-6
    assert float((og.F.cpu() - oo.F).abs().max()) <= 1e-4 * float(oo.F.abs().max())
    gr = torch.from_numpy(rng.standard_normal(tuple(oo.F.shape)).astype(np.float32))
    oo.F.backward(gr)
    og.F.backward(gr.cuda())
    assert float((fg.grad.cpu() - fo.grad).abs().max()) <= 1e-4 * float(fo.grad.abs().max())
