"""Drop-in for the `sptr` package of third_party/SparseTransformer: same names, arguments and tensor
contracts as sptr/__init__.py, sptr/functional.py, sptr/utils.py, sptr/modules.py, backed by libu2mkd_b200.so
(u2_window_pairs / u2_window_attn_fwd / u2_window_attn_bwd, include/u2mkd.h) instead of the `sptr_cuda` extension, and
without torch_scatter / torch_geometric / timm.  `u2mkd_b200.install_as_sptr()` registers it under the name `sptr`, so
core/models/sphereformer/spherical_transformer.py imports unchanged."""
import numpy as np

from .functional import (attention_step1, attention_step2, attention_step2_with_rel_pos_value, dot_prod_with_idx,
                         dot_prod_with_idx_all, precompute_all, window_attention)
from .utils import get_indices_params, grid_sample, scatter_softmax_csr, to_3d_numpy, voxel_grid


class SparseTrTensor(object):
    """sptr/__init__.py:4-33: query / key / value features + [N, 1 + 3] indices (batch first) + cached window indices."""

    def __init__(self, query_feats, query_indices, spatial_shape, batch_size, key_feats=None, value_feats=None, key_indices=None):
        self.query_feats = query_feats
        self.key_feats = key_feats
        self.value_feats = value_feats
        self.query_indices = query_indices
        self.key_indices = key_indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}

    @property
    def spatial_size(self):
        return np.prod(self.spatial_shape)

    def find_indice_params(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)


from .modules import VarLengthMultiheadSA, sparse_self_attention  # noqa: E402

__all__ = ["SparseTrTensor", "VarLengthMultiheadSA", "sparse_self_attention", "attention_step1", "attention_step2",
           "attention_step2_with_rel_pos_value", "dot_prod_with_idx", "dot_prod_with_idx_all", "precompute_all",
           "window_attention", "get_indices_params", "grid_sample", "scatter_softmax_csr", "to_3d_numpy", "voxel_grid"]
