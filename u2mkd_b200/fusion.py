"""Opt-in operator fusion on the reference's module surface.

`optimize(model)` keeps the module tree, parameter names and state_dict keys of the reference
model (core/models/build_blocks.py:21-84) and only changes how two kinds of modules execute:
  * every BatchNorm-like module (spnn.BatchNorm, nn.BatchNorm1d, [Sparse]SyncBatchNorm) runs the
    fused CUDA kernels of csrc/norm.cu (one small fp64 all-reduce per pass when synchronised);
  * a ReLU that directly follows such a module inside an nn.Sequential is folded into it;
  * a BatchNorm(+ReLU) that directly follows a sparse Conv3d inside an nn.Sequential runs as that conv's
    epilogue (ops.ConvBNReLUFn): statistics from the conv kernel, bf16 operands for the next conv and for
    the backward convs written by the BatchNorm kernels instead of separate cast passes;
  * the tail of a ResidualBlock, relu(net(x) + downsample(x)) (core/models/build_blocks.py:53-84), runs inside the
    BatchNorm epilogue of net's last conv: no separate add / ReLU passes over the block output;
  * Sequential(Linear, BatchNorm1d[, ReLU]) (the point MLPs, core/models/semantickitti/spvcnn.py:58-76) and the 1x1x1
    shortcut convs run on the same tcgen05 kernels over an identity kernel map, with the same fused BatchNorm epilogue.
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
            # [N, C, L] / [N, C, H, W] input (e.g. BatchNorm2d -> SyncBatchNorm of core/models/fusion_blocks.py:101-103): torch's
            # kernels, and the ReLU that optimize() folded into this module must still be applied
            out = super().forward(input)
            return torch.relu(out) if getattr(self, "_u2_fused_relu", False) else out
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


def _residual_forward(block):
    """forward of a ResidualBlock whose `net` ends in (Conv3d with a BatchNorm epilogue, absorbed BatchNorm).
    The block input feeds two consumers, net[0] and the shortcut.  When net[0] is a fused conv node the shortcut reads the
    input through that node's alias output, so that in backward the shortcut's gradient arrives at the node and is added
    inside its dgrad kernel instead of in a separate accumulation pass over the [N, C] gradient."""
    head, last_conv = list(block.net.children())[:-2], list(block.net.children())[-2]
    first_fused = bool(head) and _is_sparse_conv(head[0]) and getattr(head[0], "_u2_epilogue", None) is not None

    def forward(x):
        if first_fused:
            h, alias = head[0](x, want_alias=True)
            xa = type(x)(coords=x.coords, feats=alias, stride=x.stride)
            xa.cmaps, xa.kmaps = x.cmaps, x.kmaps
            shortcut = block.downsample(xa)
            rest = head[1:]
        else:
            shortcut, h, rest = block.downsample(x), x, head
        for m in rest:
            h = m(h)
        return last_conv(h, residual=shortcut.feats, relu=True)

    return forward


def _linear_bn_forward(seq):
    """forward of Sequential(Linear, BatchNorm1d[, ReLU]) (point MLPs, core/models/semantickitti/spvcnn.py:58-76): one fused
    node on the tcgen05 kernels when ops.linear_bn_relu covers it, the unchanged module chain otherwise."""
    lin, bn = list(seq.children())[:2]

    def forward(x):
        out = None
        if torch.is_tensor(x) and not getattr(bn, "_u2_absorbed", False):
            out = ops.linear_bn_relu(x, lin, bn, bool(getattr(bn, "_u2_fused_relu", False)))
        return out if out is not None else nn.Sequential.forward(seq, x)

    return forward


def _is_linear_bn(m) -> bool:
    if not isinstance(m, nn.Sequential) or "forward" in m.__dict__:
        return False
    kids = list(m.children())
    if len(kids) not in (2, 3) or not isinstance(kids[0], nn.Linear) or not isinstance(kids[1], _FusedNormMixin):
        return False
    # a third child must be the ReLU that fuse_relu folded into the BatchNorm (its forward is the identity now)
    return len(kids) == 2 or (isinstance(kids[2], nn.ReLU) and getattr(kids[2], "_u2_skip", False))


def _is_residual_block(m) -> bool:
    if type(m).__name__ != "ResidualBlock" or not all(hasattr(m, a) for a in ("net", "downsample", "relu")):
        return False
    if not isinstance(m.net, nn.Sequential) or not isinstance(m.relu, nn.ReLU) or len(m.net) < 2:
        return False
    kids = list(m.net.children())
    return (_is_sparse_conv(kids[-2]) and getattr(kids[-2], "_u2_epilogue", None) is not None
            and kids[-2]._u2_epilogue[0] is kids[-1] and not kids[-2]._u2_epilogue[1] and "forward" not in m.__dict__)


def optimize(model: nn.Module, fuse_relu: bool = True, fuse_conv_bn: bool = True, fuse_residual: bool = True) -> nn.Module:
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
            m._u2_lazy_counter = True  # num_batches_tracked of all layers: one multi-tensor add per forward
    if not getattr(model, "_u2_counter_hook", False):
        model.register_forward_hook(lambda *_: ops.flush_bn_counters())
        model._u2_counter_hook = True
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
        if fuse_residual:
            for m in model.modules():
                if _is_residual_block(m):
                    m.forward = _residual_forward(m)  # instance attribute: the class and its state_dict stay as they are
        for m in model.modules():
            if _is_linear_bn(m):
                m.forward = _linear_bn_forward(m)
    return model
