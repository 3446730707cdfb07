"""Run by tests/test_spformer_mirror_pin_cpu.py in a fresh interpreter: the SphereFormer model mirror
(u2mkd_b200/models_spformer.py — what the GPU tests and scripts/bench_spformer.py run, because the GPU box has no reference
checkout) against the UNMODIFIED reference model files (core/models/nuscenes/spvcnn_spformer.py +
core/models/sphereformer/spherical_transformer.py), both driven by the same CPU operators (oracle torchsparse / sptr
namespaces), same weights (state_dict copied reference -> mirror, strict), same batch, train mode with DropPath and dropout
off: logits and every parameter gradient.  Prints one JSON line."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import sptr_oracle, ts_oracle  # noqa: E402  (test infrastructure)

ts = ts_oracle.install_as_torchsparse()
sp = sptr_oracle.as_sptr_module()
sp.__u2_keep__ = True
sys.path.insert(0, REF)
import third_party.SparseTransformer  # noqa: E402,F401
sys.modules["third_party.SparseTransformer.sptr"] = sp

import u2mkd_b200  # noqa: E402

u2mkd_b200.install_reference_shims()
from torchpack.utils.config import configs  # noqa: E402

configs.update({"model": {"cr": 1.0, "in_channel": 4}, "data": {"num_classes": 17}})
from core.models.nuscenes.spvcnn_spformer import SPVCNN_SPFORMER as RefModel  # noqa: E402   (reference file, unchanged)

from u2mkd_b200 import models, models_spformer, scans  # noqa: E402
import copy  # noqa: E402

c, f = scans.make_batch([11, 12], "nusc", 1, 0.2)
keep = np.random.default_rng(0).permutation(c.shape[0])[:6000]
c, f = torch.from_numpy(c[keep]), torch.from_numpy(f[keep])
fam = models_spformer.build_spformer_family(models.build_family(ts), sp)
results = {}
# "builder": the argument types core/builder.py:540-553 produces from configs/nuscenes/train/spformer.yaml (float32 cubic arrays, the
# spherical window a LIST of ints, its quantisation an ndarray -> shared and scaled in place by the constructor, see the mirror);
# "arrays": everything an ndarray; "lists": the spherical sizes as float lists (copied per block)
wss = np.array([2, 2, 120])
variants = {
    "builder": dict(window_size=np.array([1.2, 1.2, 1.2], np.float32), window_size_sphere=[2, 2, 120],
                    quant_size=np.array([1.2, 1.2, 1.2], np.float32) / 24, quant_size_sphere=wss / 24),
    "arrays": dict(window_size=np.array([1.2, 1.2, 1.2], np.float32), window_size_sphere=np.array([2., 2., 120.]),
                   quant_size=np.array([1.2 / 24] * 3, np.float32), quant_size_sphere=np.array([2 / 24, 2 / 24, 120 / 24])),
    "lists": dict(window_size=np.array([1.2] * 3), window_size_sphere=[2., 2., 120.], quant_size=np.array([0.05] * 3),
                  quant_size_sphere=[1 / 12, 1 / 12, 5.]),
}
for tag, sizes in variants.items():
    kw = dict(window_size_scale=[2.0, 2.0], drop_path_rate=0.0, a=0.0125, pres=0.2, vres=0.2, **sizes)
    torch.manual_seed(0)
    ref = RefModel(**copy.deepcopy(kw))     # (both constructors scale the spherical sizes in place)
    mir = fam.SPVCNN_SPFORMER(cr=1.0, in_channel=4, num_classes=17, **copy.deepcopy(kw))
    mir.load_state_dict(ref.state_dict(), strict=True)
    qs = [[float(v) for v in b.attn.quant_size_sphere] for b in ref.transformer_blocks]
    assert qs == [[float(v) for v in b.attn.quant_size_sphere] for b in mir.transformer_blocks]
    for m in (ref, mir):
        m.train()
        m.dropout = torch.nn.Identity()
    outs, grads = [], []
    for m in (ref, mir):
        m.zero_grad()
        o = m({"lidar": ts.SparseTensor(f.clone(), c.clone())})["x_vox"]
        o.square().mean().backward()
        outs.append(o.detach())
        grads.append({k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None})
    assert set(grads[0]) == set(grads[1])
    worst = max(((float((grads[0][k] - grads[1][k]).abs().max()) / (float(grads[0][k].abs().max()) + 1e-30), k) for k in grads[0]))
    results[tag] = {"params": len(list(ref.parameters())), "state_keys": len(ref.state_dict()),
                    "logits_max_abs": float((outs[0] - outs[1]).abs().max()), "logits_ref_max": float(outs[0].abs().max()),
                    "grads_compared": len(grads[0]), "worst_grad_rel": worst[0], "worst_grad_name": worst[1],
                    "quant_size_sphere_per_block": qs,
                    "sphere_table_rows": [int(b.attn.relative_pos_query_table_sphere.shape[0]) for b in ref.transformer_blocks]}
print(json.dumps(results))
