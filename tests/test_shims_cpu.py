"""Host shims (model-file part): the stand-ins behave like the packages they replace on the calls the
reference makes, and the UNMODIFIED reference SphereFormer model files import and construct over this repo's surface."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"


def test_scatter_reductions_match_naive_loops():
    from u2mkd_b200.shims import scatter_add, scatter_max, scatter_mean
    rng = np.random.default_rng(0)
    src = torch.from_numpy(rng.standard_normal((50, 3)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, 7, size=50))
    add = torch.zeros(7, 3)
    cnt = torch.zeros(7)
    mx = torch.full((7, 3), -float("inf"))
    for i in range(50):
        add[idx[i]] += src[i]
        cnt[idx[i]] += 1
        mx[idx[i]] = torch.maximum(mx[idx[i]], src[i])
    assert torch.allclose(scatter_add(src, idx, dim=0, dim_size=7), add, atol=1e-6)
    assert torch.allclose(scatter_mean(src, idx, dim=0, dim_size=7), add / cnt.clamp_min(1)[:, None], atol=1e-6)
    got, arg = scatter_max(src, idx, dim=0, dim_size=7)
    assert torch.allclose(got, mx) and torch.equal(src.gather(0, arg), mx)
    # 1-D source, default dim, as core/models call scatter_mean(feats, inverse, dim=0)
    assert torch.allclose(scatter_mean(src[:, 0], idx, dim=0), (add / cnt.clamp_min(1)[:, None])[:, 0], atol=1e-6)


def test_drop_path_and_config():
    from u2mkd_b200.shims import Config, DropPath
    dp = DropPath(0.5)
    x = torch.ones(1000, 4)
    dp.eval()
    assert torch.equal(dp(x), x)
    dp.train()
    torch.manual_seed(0)
    y = dp(x)
    rows = y[:, 0]
    assert set(rows.unique().tolist()) <= {0.0, 2.0} and 350 < int((rows > 0).sum()) < 650 and bool((y == y[:, :1]).all())
    c = Config({"model": {"cr": 2.0}})
    c.update({"model": {"in_channel": 4}, "data": {"num_classes": 17}})
    assert c["model"]["cr"] == 2.0 and c.model.in_channel == 4 and c.data.num_classes == 17


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "core", "models")), reason="reference tree not present (GPU box)")
def test_unmodified_reference_sphereformer_model_constructs_on_this_surface():
    """core/models/nuscenes/spvcnn_spformer.py + core/models/sphereformer/spherical_transformer.py, unchanged: import
    through install_as_torchsparse() + install_reference_shims() (torchsparse, sptr, timm, torch_scatter, torchpack names
    all resolve here) and construct; the attention blocks are the reference's own module classes calling this repo's
    sparse_self_attention / get_indices_params."""
    import u2mkd_b200
    u2mkd_b200.install_as_torchsparse()
    names = u2mkd_b200.install_reference_shims()
    assert "third_party.SparseTransformer.sptr" in names
    sys.path.insert(0, REF)
    try:
        from torchpack.utils.config import configs
        configs.update({"model": {"cr": 1.0, "in_channel": 4}, "data": {"num_classes": 17}})
        from core.models.nuscenes.spvcnn_spformer import SPVCNN_SPFORMER
        import core.models.sphereformer.spherical_transformer as st
        import u2mkd_b200.sptr as our_sptr
        assert st.sparse_self_attention is our_sptr.sparse_self_attention and st.get_indices_params is our_sptr.get_indices_params
        net = SPVCNN_SPFORMER(window_size=np.array([0.3, 0.3, 0.3]), window_size_sphere=np.array([2., 2., 80.]),
                              quant_size=np.array([0.0125] * 3), quant_size_sphere=np.array([1 / 12, 1 / 12, 80 / 24]),
                              window_size_scale=[2.0, 1.5], drop_path_rate=0.3, a=0.0125, pres=0.1, vres=0.1)
        assert len(net.transformer_blocks) == 4
        attn = net.transformer_blocks[0].attn
        assert attn.relative_pos_query_table.shape == (47, 3, 1, 16) and attn.relative_pos_query_table_sphere.shape == (48, 3, 1, 16)
        assert sum(p.numel() for p in net.parameters()) > 2e7
    finally:
        sys.path.remove(REF)
