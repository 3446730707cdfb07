"""Run by tests/test_pixel_oracle_pin_cpu.py in a fresh interpreter: pins oracle/pixel_oracle.py (the checker of the CUDA
point<->pixel kernels, tests/test_gpu_pixel.py) to the REFERENCE'S OWN code on CPU:
  * core/models/fusion_blocks.py Feature_Gather / Feature_Fetch called directly (`.cuda()` = identity);
  * the multi-scale point->pixel loop of core/models/nuscenes/spvcnn_swiftnet18_spformer_tsd_full.py:448-478, which is inline
    in the student's forward: the unmodified model runs on a synthetic LiDAR + six-camera batch (oracle torchsparse / sptr
    namespaces) and forward pre-hooks capture, at each of the four stages, the loop's input (the point features, = the
    input of `learner[idx]`) and its output (`l2c_feat_map`, = the first input of `l2c_fusion_blocks[idx]`).
Prints one JSON line with the largest absolute differences."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import pixel_oracle as po, sptr_oracle, ts_oracle  # noqa: E402  (test infrastructure)

ts = ts_oracle.install_as_torchsparse()
sp = sptr_oracle.as_sptr_module()
sp.__u2_keep__ = True
sys.path.insert(0, REF)
import third_party.SparseTransformer  # noqa: E402,F401
sys.modules["third_party.SparseTransformer.sptr"] = sp
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
ts.SparseTensor.cuda = lambda self, *a, **k: self

import u2mkd_b200  # noqa: E402

u2mkd_b200.install_reference_shims()
from torchpack.utils.config import configs  # noqa: E402

os.chdir(REF)
configs.load("configs/nuscenes/train/spformer_tsd_full_ours_star.yaml", recursive=True)
configs.update(["--model.cr", "1.0", "--model.cr_t", "1.0", "--model.in_channel_t", "4", "--model.imagenet_pretrain", "None",
                "--model.teacher_pretrain", "None", "--dataset.voxel_size", "0.4"])

from core.models import fusion_blocks as fb  # noqa: E402   (reference file, unchanged)

out = {}
rng = np.random.default_rng(0)
# ---- Feature_Gather / Feature_Fetch: 2 batch elements x 6 cameras, coordinates partly outside the image, one blind camera
imgs = torch.from_numpy(rng.standard_normal((2, 6, 5, 12, 20)).astype(np.float32))
ns = [300, 170]
coords = [torch.from_numpy(rng.uniform(-1.15, 1.15, (6, n, 2)).astype(np.float32)) for n in ns]
masks = [torch.from_numpy(rng.random((6, n)) < 0.3) for n in ns]
masks[1][2] = False
out["feature_gather"] = float((fb.Feature_Gather(imgs[0], coords[0]) - po.Feature_Gather(imgs[0], coords[0])).abs().max())
out["feature_fetch"] = float((fb.Feature_Fetch(masks, coords, imgs) - po.Feature_Fetch(masks, coords, imgs)).abs().max())

# ---- the inline multi-scale loop, through the unmodified student model
from core import builder  # noqa: E402
from u2mkd_b200.shims.synthetic_nusc import SyntheticNuScenesCameras  # noqa: E402

torch.manual_seed(0)
model = builder.make_model()
stu = model.model_s
stu.train()
ds = SyntheticNuScenesCameras(voxel_size=0.4, num_train=2, num_val=2, multisweeps=2, max_points=3000, image_size=(48, 80), im_drop=0)
batch = ds["val"].collate_fn([ds["val"][0], ds["val"][1]])["feed_dict_s"]
in_mod = {"lidar": batch["lidar"], "images": batch["images"].permute(0, 1, 4, 2, 3).contiguous(),          # nusc_trainers.py:262-275
          "pixel_coordinates": batch["pixel_coordinates"], "masks": batch["masks"], "fov_mask": batch["fov_mask"].F}
in_mod["masks"][1][4] = False                                                                             # a camera that sees nothing
cap = {}
hooks = []
for idx in range(4):
    hooks.append(stu.learner[idx].register_forward_pre_hook(lambda m, a, idx=idx: cap.__setitem__(("pts", idx), a[0].detach().clone())))
    hooks.append(stu.l2c_fusion_blocks[idx].register_forward_pre_hook(
        lambda m, a, idx=idx: cap.__setitem__(("l2c", idx), (a[0].detach().clone(), tuple(a[1].shape[-2:])))))
with torch.no_grad():
    stu(in_mod)
for h in hooks:
    h.remove()
ms = []
for idx in range(4):
    ref_map, (ifh, ifw) = cap[("l2c", idx)]
    mine = po.multiscale_point2grid(cap[("pts", idx)], in_mod["pixel_coordinates"], in_mod["masks"], (ifh, ifw), 4 - idx)
    assert mine.shape == ref_map.shape, (mine.shape, ref_map.shape)
    ms.append({"stage": idx, "grid": [ifh, ifw], "channels": int(ref_map.shape[1]), "max_abs": float((mine - ref_map).abs().max()),
               "ref_max": float(ref_map.abs().max()), "nonzero": float((ref_map != 0).float().mean())})
out["multiscale"] = ms
print(json.dumps(out))
