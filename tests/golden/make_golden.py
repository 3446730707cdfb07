"""Generates tests/golden/spvcnn_ref_small.npz IN THE BUILD CONTAINER (needs /root/reference).

The reference's own, UNMODIFIED model code — core/models/semantickitti/spvcnn.py together with
core/models/utils.py and core/models/build_blocks.py — is imported from /root/reference and
run on the CPU oracle (oracle/ts_oracle.py registered as `torchsparse`, because the real
torchsparse==1.4.0 is not installable here).  Inputs, the weights' seed and the outputs are
stored, so that on the GPU box (no /root/reference) the tests can check
  (1) the oracle + our host-side mirror (u2mkd_b200/models.py) reproduce these numbers exactly,
  (2) the CUDA path reproduces them within tolerance.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import ts_oracle  # noqa: E402
from u2mkd_b200 import scans  # noqa: E402

ts_oracle.install_as_torchsparse()
from core.models.semantickitti.spvcnn import SPVCNN  # noqa: E402  (reference, unmodified)
from core.models.utils import initial_voxelize, point_to_voxel, voxel_to_point  # noqa: E402

SEED, CR, VS = 0, 0.25, 0.4


def main():
    coords, feats = scans.make_batch([11], "nusc", 1, VS)
    torch.manual_seed(SEED)
    net = SPVCNN(cr=CR, pres=VS, vres=VS, num_classes=17)
    net.dropout = torch.nn.Identity()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    x = ts_oracle.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
    out = net({"lidar": x})["x_vox"]
    target = torch.from_numpy(np.random.default_rng(SEED).integers(0, 17, size=coords.shape[0]))
    torch.nn.functional.cross_entropy(out, target).backward()

    # the glue primitives on their own (reference utils.py, unmodified)
    z = ts_oracle.PointTensor(torch.from_numpy(feats), torch.from_numpy(coords).float())
    x0 = initial_voxelize(z, VS, VS)
    z0 = voxel_to_point(x0, z)
    x1 = point_to_voxel(x0, z0)
    conv = ts_oracle.Conv3d(4, 8, 2, 2)
    torch.manual_seed(1)
    conv.reset_parameters()
    y = conv(x1)
    kmap = x1.kmaps[((1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))]

    np.savez_compressed(
        os.path.join(os.path.dirname(os.path.abspath(__file__)), "spvcnn_ref_small.npz"),
        coords=coords, feats=feats, target=target.numpy(), logits=out.detach().numpy(),
        grad_stem0=net.stem[0].kernel.grad.numpy(), grad_cls_w=net.classifier_vox[0].weight.grad.numpy(),
        grad_up3=net.vox_ups[3][0].net[0].kernel.grad.numpy(),
        state_checksum=np.array([float(sum(v.double().abs().sum() for v in state.values()))]),
        x0_coords=x0.C.numpy(), x0_feats=x0.F.detach().numpy(), z0_feats=z0.F.detach().numpy(),
        x1_feats=x1.F.detach().numpy(), idx_query_s1=z.additional_features["idx_query"][1].numpy(),
        down_coords=y.C.numpy(), down_feats=y.F.detach().numpy(), down_kernel=conv.kernel.detach().numpy(),
        down_nbmaps=kmap[0].numpy(), down_nbsizes=kmap[1].numpy(),
        meta=np.array([SEED, CR, VS]))
    print("wrote spvcnn_ref_small.npz:", coords.shape, out.shape)


if __name__ == "__main__":
    main()
