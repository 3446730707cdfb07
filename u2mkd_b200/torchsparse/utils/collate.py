"""torchsparse.utils.collate [TS v1.4.0]; core/datasets/semantic_nusc.py:363,371.
Batch index is appended as the LAST coordinate column."""
from typing import Any, List

import numpy as np
import torch

from ..tensor import SparseTensor

__all__ = ["sparse_collate", "sparse_collate_fn"]


def _as_tensor(a):
    return torch.tensor(a) if isinstance(a, np.ndarray) else a


def sparse_collate(inputs: List[SparseTensor]) -> SparseTensor:
    stride = inputs[0].stride
    coords, feats = [], []
    for b, t in enumerate(inputs):
        assert t.stride == stride, "all inputs must share one stride"
        c, f = _as_tensor(t.coords), _as_tensor(t.feats)
        assert isinstance(c, torch.Tensor) and isinstance(f, torch.Tensor)
        coords.append(torch.cat((c, torch.full((c.shape[0], 1), b, device=c.device, dtype=torch.int)), dim=1))
        feats.append(f)
    return SparseTensor(coords=torch.cat(coords, dim=0), feats=torch.cat(feats, dim=0), stride=stride)


def sparse_collate_fn(inputs: List[Any]) -> Any:
    if not isinstance(inputs[0], dict):
        return inputs
    out = {}
    for name, first in inputs[0].items():
        column = [sample[name] for sample in inputs]
        if isinstance(first, dict):
            out[name] = sparse_collate_fn(column)
        elif isinstance(first, np.ndarray):
            out[name] = torch.stack([torch.tensor(v) for v in column], dim=0)
        elif isinstance(first, torch.Tensor):
            out[name] = torch.stack(column, dim=0)
        elif isinstance(first, SparseTensor):
            out[name] = sparse_collate(column)
        else:
            out[name] = column
    return out
