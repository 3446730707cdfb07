"""Opt-in operator fusion on the reference's module surface.

`optimize(model)` keeps the module tree, parameter names and state_dict keys of the reference
model (core/models/build_blocks.py:21-84) and only changes how two kinds of modules execute:
  * every BatchNorm-like module (spnn.BatchNorm, nn.BatchNorm1d, [Sparse]SyncBatchNorm) runs the
    fused CUDA kernels of csrc/norm.cu (one small fp64 all-reduce per pass when synchronised);
  * a ReLU that directly follows such a module inside an nn.Sequential is folded into it;
  * a BatchNorm(+ReLU) that directly follows a sparse Conv3d inside an nn.Sequential runs as that conv's
    epilogue (ops.ConvBNReLUFn): statistics from the conv kernel, bf16 operands for the next conv and for
    the backward convs written by the BatchNorm kernels instead of separate cast passes.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops

_BN_TYPES = (nn.BatchNorm1d, nn.SyncBatchNorm)


def _is_sparse(x) -> bool:
    return hasattr(x, "feats") and hasattr(x, "coords")


class _FusedNormMixin:
    def forward(self, input):
        if getattr(self, "_u2_absorbed", False):  # already applied by the preceding conv (its _u2_epilogue)
            return input
        feats = input.feats if _is_sparse(input) else input
        if feats.dim() != 2:
            return super().forward(input)
        group = ops._bn_group(self)
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


def _is_sparse_conv(m) -> bool:
    from .torchsparse.nn.modules import Conv3d  # only this Conv3d honours `_u2_epilogue`
    return isinstance(m, Conv3d)


def optimize(model: nn.Module, fuse_relu: bool = True, fuse_conv_bn: bool = True) -> nn.Module:
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
    if fuse_conv_bn:
        for seq in model.modules():
            if not isinstance(seq, nn.Sequential):
                continue
            children = list(seq.children())
            for a, b in zip(children, children[1:]):
                if _is_sparse_conv(a) and isinstance(b, _FusedNormMixin) and not getattr(b, "_u2_absorbed", False) \
                        and getattr(a, "_u2_epilogue", None) is None:
                    # a tuple keeps `b` out of a's submodule registry (state_dict keys unchanged)
                    a._u2_epilogue = (b, bool(getattr(b, "_u2_fused_relu", False)))
                    b._u2_absorbed = True
    return model
