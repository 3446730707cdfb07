"""torchsparse.cat [TS v1.4.0 torchsparse/operators.py]; core/models/semantickitti/spvcnn.py:116,120,128,132."""
from typing import List

import torch

from .tensor import SparseTensor

__all__ = ["cat"]


def cat(inputs: List[SparseTensor]) -> SparseTensor:
    first = inputs[0]
    out = SparseTensor(coords=first.coords, feats=torch.cat([t.feats for t in inputs], dim=1), stride=first.stride)
    out.cmaps = first.cmaps
    out.kmaps = first.kmaps
    return out
