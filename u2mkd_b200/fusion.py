"""Opt-in operator fusion on the reference's module surface.

`optimize(model)` keeps the module tree, parameter names and state_dict keys of the reference
model (core/models/build_blocks.py:21-84) and only changes how two kinds of modules execute:
  * every BatchNorm-like module (spnn.BatchNorm, nn.BatchNorm1d, [Sparse]SyncBatchNorm) runs the
    fused CUDA kernels of csrc/norm.cu (one small fp64 all-reduce per pass when synchronised);
  * a ReLU that directly follows such a module inside an nn.Sequential is folded into it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from . import ops

_BN_TYPES = (nn.BatchNorm1d, nn.SyncBatchNorm)


def _is_sparse(x) -> bool:
    return hasattr(x, "feats") and hasattr(x, "coords")


class _FusedNormMixin:
    def forward(self, input):
        feats = input.feats if _is_sparse(input) else input
        if feats.dim() != 2:
            return super().forward(input)
        group = None
        if isinstance(self, nn.SyncBatchNorm) and self.training and dist.is_available() and dist.is_initialized():
            pg = self.process_group if self.process_group is not None else dist.group.WORLD
            if dist.get_world_size(pg) > 1:
                group = pg
        out = ops.batch_norm_relu(feats, self, relu=getattr(self, "_u2_fused_relu", False), group=group)
        if not _is_sparse(input):
            return out
        res = type(input)(coords=input.coords, feats=out, stride=input.stride)
        res.cmaps = input.cmaps
        res.kmaps = input.kmaps
        return res


_FUSED_CLASSES = {}


def _fused_class(cls):
    if cls not in _FUSED_CLASSES:
        _FUSED_CLASSES[cls] = type("U2Fused" + cls.__name__, (_FusedNormMixin, cls), {})
    return _FUSED_CLASSES[cls]


def _identity(x):
    return x


def optimize(model: nn.Module, fuse_relu: bool = True) -> nn.Module:
    """In-place; returns the model. Safe to call once, after any SyncBatchNorm conversion."""
    if fuse_relu:
        for seq in model.modules():
            if not isinstance(seq, nn.Sequential):
                continue
            children = list(seq.children())
            for a, b in zip(children, children[1:]):
                if isinstance(a, _BN_TYPES) and isinstance(b, nn.ReLU) and not getattr(b, "_u2_skip", False):
                    a._u2_fused_relu = True
                    b._u2_skip = True
                    b.forward = _identity
    for m in model.modules():
        if isinstance(m, _BN_TYPES) and not isinstance(m, _FusedNormMixin):
            m.__class__ = _fused_class(m.__class__)
    return model
