"""torchsparse.nn.utils [TS v1.4.0 nn/utils/kernel.py, nn/utils/apply.py];
core/models/utils.py:84,141."""
from typing import Callable

import numpy as np
import torch

from ..tensor import SparseTensor
from ..utils import make_ntuple

__all__ = ["get_kernel_offsets", "fapply"]


_OFFSET_CACHE = {}


def kernel_offsets_host(size, stride=1, dilation=1):
    """The offsets as a python list [K][3] (host side; ordering as get_kernel_offsets)."""
    size, stride, dilation = (make_ntuple(v, ndim=3) for v in (size, stride, dilation))
    axes = [np.arange(-size[a] // 2 + 1, size[a] // 2 + 1) * stride[a] * dilation[a] for a in range(3)]
    if int(np.prod(size)) % 2 == 1:
        return [[int(x), int(y), int(z)] for z in axes[2] for y in axes[1] for x in axes[0]]
    return [[int(x), int(y), int(z)] for x in axes[0] for y in axes[1] for z in axes[2]]


def get_kernel_offsets(size, stride=1, dilation=1, device="cpu") -> torch.Tensor:
    """int32 [K,3] neighbour offsets.  The ordering IS the weight index of Conv3d.kernel
    (SURVEY.md A.4): odd kernel volume -> x fastest, even -> z fastest.
    Device tensors are cached per (size, stride, dilation, device): the upload of a pageable host
    array is a stream synchronisation, and the reference asks for the same offsets every step.
    The returned tensor must be treated as read-only."""
    key = (make_ntuple(size, 3), make_ntuple(stride, 3), make_ntuple(dilation, 3), str(device))
    hit = _OFFSET_CACHE.get(key)
    if hit is None:
        hit = torch.tensor(np.asarray(kernel_offsets_host(size, stride, dilation)), dtype=torch.int, device=device)
        if torch.device(device).type == "cuda":
            _OFFSET_CACHE[key] = hit
    return hit


def fapply(input: SparseTensor, fn: Callable[..., torch.Tensor], *args, **kwargs) -> SparseTensor:
    out = SparseTensor(coords=input.coords, feats=fn(input.feats, *args, **kwargs), stride=input.stride)
    out.cmaps = input.cmaps
    out.kmaps = input.kmaps
    return out
