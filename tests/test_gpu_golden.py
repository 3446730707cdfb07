"""CUDA path against the committed golden fixture (outputs of the UNMODIFIED reference model
files run on the oracle, tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spvcnn_ref_small.npz")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def test_spvcnn_matches_reference_golden(cuda_lib):
    from u2mkd_b200 import models, ops
    import u2mkd_b200.torchsparse as gts
    ops.set_math("fp32")
    g = np.load(GOLDEN)
    seed, cr, vs = int(g["meta"][0]), float(g["meta"][1]), float(g["meta"][2])
    fam = models.product()
    torch.manual_seed(seed)
    net = fam.SPVCNN(cr=cr, pres=vs, vres=vs, num_classes=17)
    assert abs(float(sum(v.double().abs().sum() for v in net.state_dict().values())) - float(g["state_checksum"][0])) < 1e-6
    net.cuda()
    net.dropout = torch.nn.Identity()
    coords, feats = torch.from_numpy(g["coords"]).cuda(), torch.from_numpy(g["feats"]).cuda()
    out = net({"lidar": gts.SparseTensor(feats, coords)})["x_vox"]
    torch.nn.functional.cross_entropy(out, torch.from_numpy(g["target"]).cuda()).backward()
    assert rel(out.detach(), g["logits"]) < 1e-3
    # the fixture was written by the fp32 CPU run: whole-model gradients of two fp32 implementations differ by the
    # conditioning of 49 BatchNorm backward passes (tests/test_gpu_bench_parity.py docstring), 5e-3 is that noise level
    assert rel(net.stem[0].kernel.grad, g["grad_stem0"]) < 5e-3
    assert rel(net.vox_ups[3][0].net[0].kernel.grad, g["grad_up3"]) < 5e-3
    assert rel(net.classifier_vox[0].weight.grad, g["grad_cls_w"]) < 5e-3


def test_glue_primitives_match_reference_golden(cuda_lib):
    """initial_voxelize / voxel_to_point / point_to_voxel / strided conv + kernel map."""
    from u2mkd_b200 import models, ops
    import u2mkd_b200.torchsparse as gts
    ops.set_math("fp32")
    g = np.load(GOLDEN)
    vs = float(g["meta"][2])
    fam = models.product()
    z = gts.PointTensor(torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["coords"]).cuda().float())
    x0 = fam.initial_voxelize(z, vs, vs)
    assert np.array_equal(x0.C.cpu().numpy(), g["x0_coords"])          # voxel coordinates: bit-exact, same order
    assert np.array_equal(z.additional_features["idx_query"][1].cpu().numpy(), g["idx_query_s1"])
    assert rel(x0.F, g["x0_feats"]) < 1e-6
    z0 = fam.voxel_to_point(x0, z)
    assert rel(z0.F, g["z0_feats"]) < 1e-5
    x1 = fam.point_to_voxel(x0, z0)
    assert rel(x1.F, g["x1_feats"]) < 1e-5
    conv = gts.nn.Conv3d(4, 8, 2, 2).cuda()
    with torch.no_grad():
        conv.kernel.copy_(torch.from_numpy(g["down_kernel"]))
    y = conv(x1)
    assert np.array_equal(y.C.cpu().numpy(), g["down_coords"])
    kmap = x1.kmaps[((1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))]
    assert np.array_equal(kmap[0].cpu().numpy(), g["down_nbmaps"])     # kernel map: bit-exact
    assert np.array_equal(kmap[1].cpu().numpy().astype(np.int64), g["down_nbsizes"].astype(np.int64))
    assert rel(y.F, g["down_feats"]) < 1e-5
