"""GPU parity of the student's point <-> pixel transforms against
oracle/pixel_oracle.py, the statement-by-statement restatement of core/models/fusion_blocks.py:217-278 and
spvcnn_swiftnet18_spformer_tsd_full.py:448-494.  fp32 kernels vs the fp64 oracle, max-norm relative error <= 1e-5 (sums
of a few values per pixel); 2 batch elements x 6 cameras, points seen by none / one / several cameras."""
import numpy as np
import pytest
import torch

from oracle import pixel_oracle as po

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def make_case(seed, n_per=(5000, 3700), V=6, C=64):
    rng = np.random.default_rng(seed)
    coords, masks = [], []
    for n in n_per:
        co = torch.from_numpy(rng.uniform(-0.999, 0.999, size=(V, n, 2)).astype(np.float32))
        ma = torch.from_numpy(rng.uniform(size=(V, n)) < 0.3)
        ma[3] = False                                   # a camera that sees nothing
        coords.append(co)
        masks.append(ma)
    feats = torch.from_numpy(rng.standard_normal((sum(n_per), C)).astype(np.float32))
    return feats, coords, masks


@pytest.mark.parametrize("grid", [(28, 50), (7, 13)])
def test_point2grid_fwd_bwd(cuda_lib, grid):
    from u2mkd_b200 import pixelops
    feats, coords, masks = make_case(1)
    fo = feats.double().requires_grad_(True)
    want = po.Point2Grid(fo, [c.double() for c in coords], masks, grid)
    g = torch.from_numpy(np.random.default_rng(2).standard_normal(tuple(want.shape)))
    want.backward(g)
    fg = feats.cuda().requires_grad_(True)
    got = pixelops.Point2Grid(fg, [c.cuda() for c in coords], [m.cuda() for m in masks], grid)
    got.backward(g.float().cuda())
    assert got.shape == want.shape == (12, 64) + grid
    assert rel_err(got, want) < 1e-5 and rel_err(fg.grad, fo.grad) < 1e-5
    assert float(got[3].abs().max()) == 0.0             # the blind camera


def test_multiscale_point2grid(cuda_lib):
    from u2mkd_b200 import pixelops
    feats, coords, masks = make_case(3, C=32)
    fo = feats.double().requires_grad_(True)
    want = po.multiscale_point2grid(fo, [c.double() for c in coords], masks, (28, 50), 3)
    want.square().sum().backward()
    fg = feats.cuda().requires_grad_(True)
    got = pixelops.multiscale_point2grid(fg, [c.cuda() for c in coords], [m.cuda() for m in masks], (28, 50), 3)
    got.square().sum().backward()
    assert rel_err(got, want) < 1e-5 and rel_err(fg.grad, fo.grad) < 1e-5


def test_feature_fetch_and_gather(cuda_lib):
    from u2mkd_b200 import pixelops
    feats, coords, masks = make_case(5)
    rng = np.random.default_rng(6)
    imgs = [torch.from_numpy(rng.standard_normal((6, 48, 23, 41)).astype(np.float32)) for _ in range(2)]
    io = [i.double().requires_grad_(True) for i in imgs]
    want = po.Feature_Fetch(masks, [c.double() for c in coords], io)
    g = torch.from_numpy(rng.standard_normal(tuple(want.shape)))
    want.backward(g)
    ig = [i.cuda().requires_grad_(True) for i in imgs]
    got = pixelops.Feature_Fetch([m.cuda() for m in masks], [c.cuda() for c in coords], ig)
    got.backward(g.float().cuda())
    assert got.shape == want.shape == (8700, 48)
    assert rel_err(got, want) < 1e-5
    for a, b in zip(ig, io):
        assert rel_err(a.grad, b.grad) < 1e-5
    seen = torch.cat([m.any(0) for m in masks])
    assert float(got[~seen.cuda()].abs().max()) == 0.0  # points no camera sees
    # Feature_Gather: every point in every camera, coordinates slightly outside the image included (zero padding)
    xy = torch.from_numpy(rng.uniform(-1.1, 1.1, size=(6, 900, 2)).astype(np.float32))
    w = po.Feature_Gather(imgs[0].double(), xy.double())
    gg = pixelops.Feature_Gather(imgs[0].cuda(), xy.cuda())
    assert gg.shape == w.shape == (6, 48, 900) and rel_err(gg, w) < 1e-5
