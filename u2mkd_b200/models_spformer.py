"""Host mirror of the SphereFormer teacher backbone: core/models/sphereformer/spherical_transformer.py
(cart2sphere :31-36, exponential_split :39-66, SparseMultiheadSASphereConcat :70-283, SphereFormer :286-347, Mlp :10-28)
and core/models/nuscenes/spvcnn_spformer.py (SPVCNN_SPFORMER).  Same module tree and state_dict keys as the reference
files; it exists because /root/reference is absent on the GPU box (the unmodified files import and construct on this
surface, tests/test_shims_cpu.py).  Instantiated over a torchsparse-like family (u2mkd_b200.models.build_family) and an
sptr-like module: the product binds u2mkd_b200.torchsparse + u2mkd_b200.sptr, the tests bind the CPU oracles."""
from functools import partial
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from .shims import DropPath


def cart2sphere(xyz):
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    theta = (torch.atan2(y, x) + np.pi) * 180 / np.pi
    beta = torch.atan2(torch.sqrt(x ** 2 + y ** 2), z) * 180 / np.pi
    r = torch.sqrt(x ** 2 + y ** 2 + z ** 2)
    return torch.stack([theta, beta, r], -1)


def exponential_split(xyz, index_0, index_1, relative_position_index, a=0.05 * 0.25):
    """Radial index of the spherical branch: bins that double in length every second step (spherical_transformer.py:39-66)."""
    r = xyz[:, 2]
    rel_pos = r[index_0.long()] - r[index_1.long()]
    rel_pos_abs = rel_pos.abs()
    flag_float = (rel_pos >= 0).to(rel_pos.dtype)
    idx = 2 * torch.floor(torch.log((rel_pos_abs + 2 * a) / a) / np.log(2)) - 2
    idx = idx + ((3 * (2 ** (idx // 2)) - 2) * a <= rel_pos_abs).to(rel_pos.dtype)
    idx = idx * (2 * flag_float - 1) + (flag_float - 1)
    relative_position_index[:, 2] = idx.long() + 24
    return relative_position_index


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop, inplace=True)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


def build_spformer_family(fam, sptr) -> SimpleNamespace:
    to_3d_numpy, SparseTrTensor = sptr.to_3d_numpy, sptr.SparseTrTensor
    sparse_self_attention, get_indices_params = sptr.sparse_self_attention, sptr.get_indices_params
    ts = fam.ts
    PointTensor = ts.PointTensor

    class SparseMultiheadSASphereConcat(nn.Module):
        """Half of the heads attend inside cubic windows, the other half inside spherical (theta, beta, r) windows
        (spherical_transformer.py:70-283), contextual relative position encoding on both."""

        def __init__(self, embed_dim, num_heads, indice_key, window_size, window_size_sphere, shift_win=False, pe_type='none',
                     dropout=0., qk_scale=None, qkv_bias=True, algo='native', **kwargs):
            super().__init__()
            assert pe_type == 'contextual', "the U2MKD models use the contextual encoding only"
            self.embed_dim, self.num_heads, self.indice_key, self.shift_win, self.pe_type = embed_dim, num_heads, indice_key, shift_win, pe_type
            head_dim = embed_dim // num_heads
            self.scale = qk_scale or head_dim ** -0.5
            self.window_size = to_3d_numpy(window_size)
            self.window_size_sphere = to_3d_numpy(window_size_sphere)
            self.rel_query, self.rel_key, self.rel_value = kwargs['rel_query'], kwargs['rel_key'], kwargs['rel_value']
            self.quant_size = to_3d_numpy(kwargs['quant_size'])
            self.quant_size_sphere = to_3d_numpy(kwargs['quant_size_sphere'])
            self.a = kwargs['a']
            assert self.rel_query and self.rel_key and self.rel_value
            qgl = int((self.window_size[0] + 1e-4) / self.quant_size[0])
            assert qgl == int((self.window_size[1] + 1e-4) / self.quant_size[1])
            self.num_heads_brc1 = num_heads // 2
            for name in ("query", "key", "value"):
                t = nn.Parameter(torch.zeros(2 * qgl - 1, 3, self.num_heads_brc1, head_dim))
                nn.init.trunc_normal_(t, std=.02)
                setattr(self, f"relative_pos_{name}_table", t)
            self.quant_grid_length = qgl
            qgls = int((self.window_size_sphere[0] + 1e-4) / self.quant_size_sphere[0])
            assert qgls == int((self.window_size_sphere[1] + 1e-4) / self.quant_size_sphere[1])
            for name in ("query", "key", "value"):
                t = nn.Parameter(torch.zeros(2 * qgls, 3, num_heads - self.num_heads_brc1, head_dim))
                nn.init.trunc_normal_(t, std=.02)
                setattr(self, f"relative_pos_{name}_table_sphere", t)
            self.quant_grid_length_sphere = qgls
            self.qkv = nn.Linear(embed_dim, embed_dim * 3, bias=qkv_bias)
            self.attn_drop = nn.Dropout(dropout, inplace=True)
            self.proj = nn.Linear(embed_dim, embed_dim)
            self.proj_drop = nn.Dropout(dropout, inplace=True)

        def forward(self, sptr_tensor):
            query = sptr_tensor.query_feats
            assert sptr_tensor.key_feats is None and sptr_tensor.value_feats is None
            xyz = sptr_tensor.query_indices[:, 1:]
            batch = sptr_tensor.query_indices[:, 0]
            N, C = query.shape
            qkv = self.qkv(query).reshape(N, 3, self.num_heads, C // self.num_heads).permute(1, 0, 2, 3).contiguous()
            query, key, value = qkv[0], qkv[1], qkv[2]
            query = query * self.scale
            xyz_sphere = cart2sphere(xyz)
            params = sptr_tensor.find_indice_params(self.indice_key)
            if params is None:
                cub = get_indices_params(xyz, batch, self.window_size, self.shift_win)
                sph = get_indices_params(xyz_sphere, batch, self.window_size_sphere, self.shift_win)
                sptr_tensor.indice_dict[self.indice_key] = (cub, sph)
            else:
                cub, sph = params
            h1 = self.num_heads_brc1
            outs = []
            for (i0, i0o, n_max, i1, i1o, sort_idx), coords, sl, ws, qs, qgl, sfx, split in (
                    (cub, xyz, slice(0, h1), self.window_size, self.quant_size, self.quant_grid_length, "", None),
                    (sph, xyz_sphere, slice(h1, None), self.window_size_sphere, self.quant_size_sphere,
                     self.quant_grid_length_sphere, "_sphere", partial(exponential_split, a=self.a))):
                outs.append(sparse_self_attention(
                    query=query[:, sl].contiguous().to(coords.dtype), key=key[:, sl].contiguous().to(coords.dtype),
                    value=value[:, sl].contiguous().to(coords.dtype), xyz=coords, index_0=i0.int(), index_0_offsets=i0o.int(),
                    n_max=n_max, index_1=i1.int(), index_1_offsets=i1o.int(), sort_idx=sort_idx, window_size=ws,
                    shift_win=self.shift_win, pe_type=self.pe_type, rel_query=True, rel_key=True, rel_value=True, quant_size=qs,
                    quant_grid_length=qgl,
                    relative_pos_query_table=getattr(self, "relative_pos_query_table" + sfx).to(coords.dtype),
                    relative_pos_key_table=getattr(self, "relative_pos_key_table" + sfx).to(coords.dtype),
                    relative_pos_value_table=getattr(self, "relative_pos_value_table" + sfx).to(coords.dtype), split_func=split))
            x = torch.cat(outs, 1).view(N, C).to(self.proj.weight.dtype)
            x = self.proj_drop(self.proj(x))
            return SparseTrTensor(x, sptr_tensor.query_indices, sptr_tensor.spatial_shape, sptr_tensor.batch_size)

    class SphereFormer(nn.Module):
        def __init__(self, dim, num_heads, window_size, window_size_sphere, quant_size, quant_size_sphere, indice_key,
                     pe_type='contextual', rel_query=True, rel_key=False, rel_value=False, drop_path=0.0, mlp_ratio=4.0,
                     qkv_bias=True, qk_scale=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, a=0.05 * 0.25):
            super().__init__()
            self.window_size = window_size
            self.norm1 = norm_layer(dim)
            self.attn = SparseMultiheadSASphereConcat(dim, num_heads=num_heads, indice_key=indice_key, window_size=window_size,
                                                      window_size_sphere=window_size_sphere, pe_type=pe_type, quant_size=quant_size,
                                                      quant_size_sphere=quant_size_sphere, rel_query=rel_query, rel_key=rel_key,
                                                      rel_value=rel_value, qkv_bias=qkv_bias, qk_scale=qk_scale, a=a)
            self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
            self.norm2 = norm_layer(dim)
            self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)

        def forward(self, feats, xyz, batch):
            short_cut = feats
            feats = self.norm1(feats)
            t = SparseTrTensor(feats, torch.cat([batch[:, None], xyz], -1), spatial_shape=None, batch_size=None)
            feats = self.attn(t).query_feats
            feats = short_cut + self.drop_path(feats)
            return feats + self.drop_path(self.mlp(self.norm2(feats)))

    class SPVCNN_SPFORMER(fam.SPVCNN):
        """core/models/nuscenes/spvcnn_spformer.py: SPVCNN with a SphereFormer block after every down stage (:145-149); the
        window / quantisation sizes grow by window_size_scale from stage to stage (:76-83).  cr / in_channel / num_classes
        come as keyword arguments here (the reference reads them from torchpack's global `configs`)."""

        def __init__(self, window_size, window_size_sphere, quant_size, quant_size_sphere, window_size_scale, drop_path_rate, a,
                     pres, vres, **kwargs):
            super().__init__(pres=pres, vres=vres, **kwargs)
            cs = [int(kwargs.get("cr", 1.0) * c) for c in (32, 32, 64, 128, 256, 256, 128, 96, 96)]
            # spvcnn_spformer.py:51-83 statement by statement, on the caller's own objects and dtypes — this matters: the cubic
            # sizes are REBOUND per stage (`x = x * s`, new arrays) while the spherical ones are updated IN PLACE
            # (`x[0] = x[0] * s`), and sptr's to_3d_numpy copies a list but returns an ndarray as is.  With builder.make_model's
            # arguments (core/builder.py:540-553: window_size_sphere a list, quant_size_sphere an ndarray) every block therefore
            # ends up sharing ONE quant_size_sphere array holding the final, 16 x scaled values at forward time, while its
            # table length was fixed from the value at construction.  Reproduced as is (pinned against the unmodified model
            # files on CPU, tests/test_spformer_mirror_pin_cpu.py).
            self.window_size, self.window_size_sphere = window_size, window_size_sphere
            self.quant_size, self.quant_size_sphere = quant_size, quant_size_sphere
            dpr = [x.item() for x in torch.linspace(0, drop_path_rate, 7)]
            head_dim = 16
            self.transformer_blocks = nn.ModuleList()
            for idx in range(1, 5):
                self.transformer_blocks.append(SphereFormer(
                    cs[idx], cs[idx] // head_dim, self.window_size, self.window_size_sphere, self.quant_size,
                    self.quant_size_sphere, indice_key='sphereformer{}'.format(idx + 1), rel_query=True, rel_key=True,
                    rel_value=True, drop_path=dpr[idx], a=a))
                sc, ss = window_size_scale
                self.window_size = self.window_size * sc
                self.quant_size = self.quant_size * sc
                self.window_size_sphere[0] = self.window_size_sphere[0] * ss
                self.window_size_sphere[1] = self.window_size_sphere[1] * ss
                self.quant_size_sphere[0] = self.quant_size_sphere[0] * ss
                self.quant_size_sphere[1] = self.quant_size_sphere[1] * ss
            for m in self.modules():
                if isinstance(m, nn.BatchNorm1d):
                    nn.init.constant_(m.weight, 1)
                    nn.init.constant_(m.bias, 0)

        def forward(self, in_mod):
            x = in_mod["lidar"]
            z = PointTensor(x.F, x.C.float())
            x0 = fam.initial_voxelize(z, self.pres, self.vres)
            zz = PointTensor(x0.F, x0.C.float())
            x0 = self.stem(x0)
            z0 = fam.voxel_to_point(x0, z, nearest=False)
            feats = [fam.point_to_voxel(x0, z0)]
            for idx, down in enumerate(self.vox_downs):
                vox_out = down(feats[idx])
                tmp_p = fam.point_to_voxel(vox_out, zz)
                coord_xyz, batch = tmp_p.F[:, :3], tmp_p.C[:, 3]
                vox_out.F = self.transformer_blocks[idx](vox_out.F, coord_xyz, batch)
                feats.append(vox_out)
            x1, x2, x3, x4 = feats[1:]
            z1 = fam.voxel_to_point(x4, z0)
            z1.F = z1.F + self.point_transforms[0](z0.F)
            y1 = fam.point_to_voxel(x4, z1)
            y1.F = self.dropout(y1.F)
            y1 = self._up(0, y1, x3)
            y2 = self._up(1, y1, x2)
            z2 = fam.voxel_to_point(y2, z1)
            z2.F = z2.F + self.point_transforms[1](z1.F)
            y3 = fam.point_to_voxel(y2, z2)
            y3.F = self.dropout(y3.F)
            y3 = self._up(2, y3, x1)
            y4 = self._up(3, y3, x0)
            z3 = fam.voxel_to_point(y4, z2)
            z3.F = z3.F + self.point_transforms[2](z2.F)
            return {"x_vox": self.classifier_vox(z3.F)}

    return SimpleNamespace(SparseMultiheadSASphereConcat=SparseMultiheadSASphereConcat, SphereFormer=SphereFormer,
                           SPVCNN_SPFORMER=SPVCNN_SPFORMER, cart2sphere=cart2sphere, exponential_split=exponential_split)


_product = None


def product() -> SimpleNamespace:
    """The SphereFormer family bound to the CUDA path (u2mkd_b200.torchsparse + u2mkd_b200.sptr)."""
    global _product
    if _product is None:
        from . import models, sptr
        _product = build_spformer_family(models.product(), sptr)
    return _product
