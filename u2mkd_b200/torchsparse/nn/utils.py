"""torchsparse.nn.utils [TS v1.4.0 nn/utils/kernel.py, nn/utils/apply.py];
core/models/utils.py:84,141."""
from typing import Callable

import numpy as np
import torch

from ..tensor import SparseTensor
from ..utils import make_ntuple

__all__ = ["get_kernel_offsets", "fapply"]


def get_kernel_offsets(size, stride=1, dilation=1, device="cpu") -> torch.Tensor:
    """int32 [K,3] neighbour offsets.  The ordering IS the weight index of Conv3d.kernel
    (SURVEY.md A.4): odd kernel volume -> x fastest, even -> z fastest."""
    size, stride, dilation = (make_ntuple(v, ndim=3) for v in (size, stride, dilation))
    axes = [np.arange(-size[a] // 2 + 1, size[a] // 2 + 1) * stride[a] * dilation[a] for a in range(3)]
    if int(np.prod(size)) % 2 == 1:
        grid = [[x, y, z] for z in axes[2] for y in axes[1] for x in axes[0]]
    else:
        grid = [[x, y, z] for x in axes[0] for y in axes[1] for z in axes[2]]
    return torch.tensor(np.asarray(grid), dtype=torch.int, device=device)


def fapply(input: SparseTensor, fn: Callable[..., torch.Tensor], *args, **kwargs) -> SparseTensor:
    out = SparseTensor(coords=input.coords, feats=fn(input.feats, *args, **kwargs), stride=input.stride)
    out.cmaps = input.cmaps
    out.kmaps = input.kmaps
    return out
