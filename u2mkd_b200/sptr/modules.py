"""sptr.modules (third_party/SparseTransformer/sptr/modules.py): sparse_self_attention and VarLengthMultiheadSA."""
import numpy as np
import torch
import torch.nn as nn

from . import SparseTrTensor
from .functional import (attention_step1, attention_step2, attention_step2_with_rel_pos_value, dot_prod_with_idx_all,
                         window_attention)
from .utils import get_indices_params, scatter_softmax_csr, to_3d_numpy
from .._lib import lib


def sparse_self_attention(query, key, value, xyz, index_0, index_0_offsets, n_max, index_1, index_1_offsets, sort_idx,
                          window_size, shift_win, pe_type='none', rel_query=False, rel_key=False, rel_value=False,
                          quant_size=None, quant_grid_length=None, relative_pos_query_table=None,
                          relative_pos_key_table=None, relative_pos_value_table=None, split_func=None):
    """sptr/modules.py:11-62, same arguments and result.  The relative position index is computed exactly as there
    (:40-48); scores, softmax and the weighted sum then run as ONE fused kernel (functional.window_attention) when the
    window boundaries came with index_0_offsets (get_indices_params / precompute_all attach them) and the head size /
    table length are covered — otherwise as the reference's three steps."""
    query = query[sort_idx]
    key = key[sort_idx]
    value = value[sort_idx]
    xyz_ctg = xyz[sort_idx]
    contextual = pe_type == 'contextual' and rel_query and rel_key
    relative_position_index = None
    if contextual:
        window_size = torch.from_numpy(np.asarray(window_size)).float().to(xyz.device)
        shift_size = 1 / 2 * window_size if shift_win else 0.0
        xyz_quant = (xyz_ctg - xyz_ctg.min(0)[0] + shift_size) % window_size
        xyz_quant = torch.div(xyz_quant, torch.from_numpy(np.asarray(quant_size)).float().to(xyz.device), rounding_mode='floor')
        relative_position = xyz_quant[index_0.long()] - xyz_quant[index_1.long()]  # [M, 3]
        relative_position_index = relative_position + quant_grid_length - 1
        if split_func:
            relative_position_index = split_func(xyz_ctg, index_0, index_1, relative_position_index.clone())
            relative_position_index = torch.clamp(relative_position_index, 0, 2 * quant_grid_length - 1)
        relative_position_index = relative_position_index.int()

    windows = getattr(index_0_offsets, "_u2_windows", None)
    with_value_table = pe_type == 'contextual' and rel_value
    L = relative_pos_query_table.shape[0] if contextual else 0
    fusable = (windows is not None and query.is_cuda and contextual == with_value_table
               and lib().u2_window_attn_supported(query.shape[2], L))
    if fusable:
        win_off, sq_off, n_windows = windows
        if contextual:
            x = window_attention(query, key, value, win_off, sq_off, n_windows, relative_position_index,
                                 relative_pos_query_table, relative_pos_key_table, relative_pos_value_table)
        else:
            x = window_attention(query, key, value, win_off, sq_off, n_windows)
    else:
        if contextual:
            attn_flat = dot_prod_with_idx_all(query, index_0, index_0_offsets, key, index_1, index_1_offsets,
                                              relative_pos_query_table, relative_pos_key_table, relative_position_index, n_max)
        else:
            attn_flat = attention_step1(query, key, index_0, index_0_offsets, index_1, index_1_offsets, n_max)
        softmax_attn_flat = scatter_softmax_csr(src=attn_flat, indptr=index_0_offsets.long(), dim=0)  # [M, num_heads]
        if with_value_table:
            x = attention_step2_with_rel_pos_value(softmax_attn_flat, value, index_0, index_0_offsets, n_max, index_1,
                                                   index_1_offsets, relative_pos_value_table, relative_position_index)
        else:
            x = attention_step2(softmax_attn_flat, value, index_0, index_0_offsets, index_1, index_1_offsets, n_max)
    out = torch.empty_like(x)
    out[sort_idx] = x
    return out


class VarLengthMultiheadSA(nn.Module):
    """sptr/modules.py:65-200: multi-head self-attention over variable-length windows ('none' and 'contextual' position
    encodings; the sine / fourier encodings of sptr/position_embedding.py are not used by any U2MKD model)."""

    def __init__(self, embed_dim, num_heads, indice_key, window_size, shift_win=False, pe_type='none', dropout=0., qk_scale=None,
                 qkv_bias=True, algo='native', **kwargs):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.indice_key = indice_key
        self.shift_win = shift_win
        self.pe_type = pe_type
        head_dim = embed_dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.window_size = to_3d_numpy(window_size)
        if pe_type == 'contextual':
            self.rel_query, self.rel_key, self.rel_value = kwargs['rel_query'], kwargs['rel_key'], kwargs['rel_value']
            quant_size = kwargs['quant_size']
            self.quant_size = to_3d_numpy(quant_size)
            window_size3, quant_size3 = to_3d_numpy(window_size), self.quant_size
            quant_grid_length = int((window_size3[0] + 1e-4) / quant_size3[0])
            assert int((window_size3[0] + 1e-4) / quant_size3[0]) == int((window_size3[1] + 1e-4) / quant_size3[1])
            assert self.rel_query and self.rel_key and self.rel_value
            for name in ("query", "key", "value"):
                t = nn.Parameter(torch.zeros(2 * quant_grid_length - 1, 3, num_heads, head_dim))
                nn.init.trunc_normal_(t, std=.02)
                setattr(self, f"relative_pos_{name}_table", t)
            self.quant_grid_length = quant_grid_length
        elif pe_type != 'none':
            raise ValueError(f"pe_type {pe_type!r}: only 'none' and 'contextual' are built (no U2MKD model uses sine / fourier)")
        self.q = nn.Linear(embed_dim, embed_dim, bias=qkv_bias)
        self.k = nn.Linear(embed_dim, embed_dim, bias=qkv_bias)
        self.v = nn.Linear(embed_dim, embed_dim, bias=qkv_bias)
        self.attn_drop = nn.Dropout(dropout, inplace=True)
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.proj_drop = nn.Dropout(dropout, inplace=True)

    def forward(self, sptr_tensor: SparseTrTensor):
        query, key, value = sptr_tensor.query_feats, sptr_tensor.key_feats, sptr_tensor.value_feats
        if key is None:
            key = query.clone()
        if value is None:
            value = query.clone()
        xyz = sptr_tensor.query_indices[:, 1:]
        batch = sptr_tensor.query_indices[:, 0]
        assert xyz.shape[1] == 3
        N, C = query.shape
        query = self.q(query).reshape(N, self.num_heads, C // self.num_heads)
        key = self.k(key).reshape(N, self.num_heads, C // self.num_heads)
        value = self.v(value).reshape(N, self.num_heads, C // self.num_heads)
        query = query * self.scale
        index_params = sptr_tensor.find_indice_params(self.indice_key)
        if index_params is None:
            index_0, index_0_offsets, n_max, index_1, index_1_offsets, sort_idx = get_indices_params(xyz, batch, self.window_size, self.shift_win)
            sptr_tensor.indice_dict[self.indice_key] = (index_0, index_0_offsets, n_max, index_1, index_1_offsets, sort_idx,
                                                        self.window_size, self.shift_win)
        else:
            index_0, index_0_offsets, n_max, index_1, index_1_offsets, sort_idx, window_size, shift_win = index_params
            assert (window_size == self.window_size).all() and (shift_win == self.shift_win), \
                "window_size and shift_win must be the same for sptr_tensors with the same indice_key: {}".format(self.indice_key)
        kwargs = {"query": query.float(), "key": key.float(), "value": value.float(), "xyz": xyz.float(),
                  "index_0": index_0.int(), "index_0_offsets": index_0_offsets.int(), "n_max": n_max, "index_1": index_1.int(),
                  "index_1_offsets": index_1_offsets.int(), "sort_idx": sort_idx, "window_size": self.window_size,
                  "shift_win": self.shift_win, "pe_type": self.pe_type}
        if self.pe_type == 'contextual':
            kwargs.update({"rel_query": self.rel_query, "rel_key": self.rel_key, "rel_value": self.rel_value,
                           "quant_size": self.quant_size, "quant_grid_length": self.quant_grid_length,
                           "relative_pos_query_table": self.relative_pos_query_table.float(),
                           "relative_pos_key_table": self.relative_pos_key_table.float(),
                           "relative_pos_value_table": self.relative_pos_value_table.float()})
        x = sparse_self_attention(**kwargs)
        x = x.view(N, C)
        x = self.proj(x)
        x = self.proj_drop(x)
        return SparseTrTensor(x, sptr_tensor.query_indices, sptr_tensor.spatial_shape, sptr_tensor.batch_size)
