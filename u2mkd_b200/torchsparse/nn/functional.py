"""torchsparse.nn.functional: sphash / sphashquery / spcount / spvoxelize / spdevoxelize /
calc_ti_weights / spdownsample / conv3d, all backed by libu2mkd_b200.so (include/u2mkd.h).
Call sites: core/models/utils.py:19-26,43-58,84-99,133-135; conv3d through spnn.Conv3d."""
import functools

import torch

from ... import ops
from ...ops import calc_ti_weights, coord_query, spcount, spdevoxelize, sphash, sphashquery, spvoxelize, unique_voxelize
from ..tensor import SparseTensor
from ..utils import make_ntuple
from .utils import get_kernel_offsets, kernel_offsets_host

__all__ = ["sphash", "sphashquery", "spcount", "spvoxelize", "spdevoxelize", "calc_ti_weights", "spdownsample",
           "conv3d", "unique_voxelize", "coord_query", "build_map_for", "prebuild_maps"]

prebuild_maps = ops.prebuild_maps


def spdownsample(coords: torch.Tensor, stride=2, kernel_size=2, tensor_stride=1) -> torch.Tensor:
    """Output coordinates of a strided conv, sorted by (b,x,y,z)."""
    stride, kernel_size, tensor_stride = (make_ntuple(v, ndim=3) for v in (stride, kernel_size, tensor_stride))
    sample_stride = tuple(stride[a] * tensor_stride[a] for a in range(3))
    if all(stride[a] in (1, kernel_size[a]) for a in range(3)):
        return ops.downsample_coords(coords, sample_stride)
    # general case (no U2MKD model uses it): expand by the kernel offsets, keep the
    # candidates that sit on the output lattice, then the same sorted-unique kernel
    offsets = get_kernel_offsets(kernel_size, tensor_stride, device=coords.device)
    kv = offsets.size(0)
    ss = torch.tensor(sample_stride, dtype=torch.int, device=coords.device).unsqueeze(0)
    cmin = torch.min(coords[:, :3], dim=0, keepdim=True).values
    xyz = (coords[:, :3].unsqueeze(1) + offsets.unsqueeze(0)).view(-1, 3)
    b = coords[:, 3:].repeat(1, kv).view(-1, 1)
    keep = torch.all((xyz % ss == 0) & (xyz >= cmin), dim=1)
    cand = torch.cat([xyz, b], dim=1)[keep].contiguous()
    return ops.downsample_coords(cand, (1, 1, 1))


def build_map_for(cmaps: dict, kmaps: dict, key, in_coords=None):
    """Kernel map `key` = (tensor_stride, kernel_size, stride, dilation) over the coordinate sets in `cmaps`: output
    coordinates by spdownsample for a strided conv (registered under the output stride unless that set already exists),
    then the neighbour tables.  Shared by F.conv3d (cache miss) and ops.prebuild_maps (start of the forward pass)."""
    in_stride, kernel_size, stride, dilation = key
    coords_in = cmaps[in_stride] if in_coords is None else in_coords
    offsets = get_kernel_offsets(kernel_size, stride=in_stride, device=coords_in.device)
    coords_out = coords_in
    if any(s > 1 for s in stride):
        out_stride = tuple(in_stride[a] * stride[a] for a in range(3))
        coords_out = cmaps.get(out_stride)
        if coords_out is None:
            coords_out = spdownsample(coords_in, stride, kernel_size, in_stride)
            cmaps[out_stride] = coords_out
    kmap = ops.build_kernel_map(coords_in, coords_out, offsets, kernel_offsets_host(kernel_size, in_stride))
    kmap.plan_key = key
    ops._plan_note(key)
    kmaps[key] = kmap
    return kmap


def conv3d(input: SparseTensor, weight: torch.Tensor, kernel_size, bias=None, stride=1, dilation=1,
           transposed: bool = False, epilogue=None, residual=None, want_alias: bool = False) -> SparseTensor:
    """F.conv3d of torchsparse v1.4.0: kernel-map lookup/build, then
    one fused gather-GEMM kernel (ops.ConvolutionFn) instead of K gather/mm/scatter rounds.
    `epilogue=(bn_module, relu)` (set by u2mkd_b200.fusion.optimize, not part of the torchsparse signature)
    applies that BatchNorm(+ReLU) to the result inside the same autograd node; `residual` (a feature matrix,
    only with an epilogue) is added between the BatchNorm and the ReLU (ResidualBlock tail).  want_alias (fusion only):
    returns (output, alias) with alias = the input feature matrix routed through the fused conv node, see
    ops.sparse_conv_bn_relu."""
    assert residual is None or epilogue is not None
    alias = input.feats
    feats, coords = input.feats, input.coords
    kernel_size, stride, dilation = (make_ntuple(v, ndim=3) for v in (kernel_size, stride, dilation))
    unit = (1, 1, 1)
    if kernel_size == unit and stride == unit and dilation == unit:
        out_stride = input.stride
        if feats.is_cuda and feats.dim() == 2 and weight.dim() == 2 and feats.shape[0] > 1 and ops.dense_tc_supported(*weight.shape):
            # bf16 mode: the dense layer runs on the tcgen05 conv kernels over an identity kernel map (no cuBLAS), with the
            # BatchNorm epilogue fused like every other conv
            kmap = ops.identity_kernel_map(feats.shape[0], feats.device)
            if epilogue is not None and bias is None:
                res = ops.sparse_conv_bn_relu(feats, weight.unsqueeze(0), kmap, False, *epilogue, residual=residual,
                                              want_alias=want_alias)
                feats, alias = res if want_alias else (res, alias)
                epilogue = None
            else:
                feats = ops.sparse_conv(feats, weight.unsqueeze(0), kmap, transposed=False)
        else:
            feats = feats.matmul(weight)
    elif not transposed:
        out_stride = tuple(input.stride[a] * stride[a] for a in range(3))
        key = (input.stride, kernel_size, stride, dilation)
        kmap = input.kmaps.get(key)
        if kmap is None:
            input.cmaps.setdefault(input.stride, input.coords)
            kmap = build_map_for(input.cmaps, input.kmaps, key, in_coords=input.coords)
        if any(s > 1 for s in stride):
            coords = input.cmaps[out_stride]  # upstream skips this on a cache hit
        if epilogue is not None and bias is None:
            res = ops.sparse_conv_bn_relu(feats, weight, kmap, False, *epilogue, residual=residual, want_alias=want_alias)
            feats, alias = res if want_alias else (res, alias)
            epilogue = None
        else:
            feats = ops.sparse_conv(feats, weight, kmap, transposed=False)
    else:
        out_stride = tuple(input.stride[a] // stride[a] for a in range(3))
        kmap = input.kmaps[(out_stride, kernel_size, stride, dilation)]
        if epilogue is not None and bias is None:
            res = ops.sparse_conv_bn_relu(feats, weight, kmap, True, *epilogue, residual=residual, want_alias=want_alias)
            feats, alias = res if want_alias else (res, alias)
            epilogue = None
        else:
            feats = ops.sparse_conv(feats, weight, kmap, transposed=True)
        coords = input.cmaps[out_stride]
    if bias is not None:
        feats = feats + bias
    if epilogue is not None:  # 1x1x1 kernels and biased convs: the separate BatchNorm kernels
        if residual is None:
            feats = ops.batch_norm_relu(feats, epilogue[0], epilogue[1], ops._bn_group(epilogue[0]))
        else:
            feats = ops.batch_norm_relu(feats, epilogue[0], False, ops._bn_group(epilogue[0])) + residual
            feats = torch.relu_(feats) if epilogue[1] else feats
    output = SparseTensor(coords=coords, feats=feats, stride=out_stride)
    output.cmaps = input.cmaps
    output.cmaps.setdefault(output.stride, output.coords)
    output.kmaps = input.kmaps
    return (output, alias) if want_alias else output
