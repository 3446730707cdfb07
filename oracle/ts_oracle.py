"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the torchsparse v1.4.0 operator surface that U2MKD's LiDAR
point-voxel path calls.  torchsparse is pinned by
the reference at /root/reference/README.md:44-48 (mit-han-lab/torchsparse@v1.4.0) and
is NOT vendored, so its published algorithm is restated here; every function cites the
reference call site it serves and the upstream file it follows.

PARITY UNPINNED: the reference has no golden vectors / KATs / fixtures for this path
(SURVEY.md §4, §8(c)).  Pins are our own: dense-equivalence KATs against F.conv3d /
F.conv_transpose3d / F.grid_sample / avg_pool3d, brute-force dict kernel maps and fp64
gradcheck (tests/test_oracle_kat.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  Integer/hash ops and gather/scatter run in C
(oracle/ts_cpu.c, OpenMP); the per-offset GEMM is torch.mm, exactly like the reference
CPU backend's torch::mm_out.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
import types
from itertools import repeat
from typing import List, Tuple, Union

import numpy as np
import torch
from torch import nn
from torch.autograd import Function

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/ts_cpu.c -> oracle/_build/libts_cpu.so (gcc, OpenMP)."""
    so = os.path.join(_HERE, "_build", "libts_cpu.so")
    src = os.path.join(_HERE, "ts_cpu.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared",
                               "-o", so, src])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def _i64(v):
    return ctypes.c_int64(int(v))


def _suf(t: torch.Tensor) -> str:
    if t.dtype == torch.float32:
        return "f32"
    if t.dtype == torch.float64:
        return "f64"
    raise TypeError(f"oracle supports fp32/fp64, got {t.dtype}")


# --------------------------------------------------------------------------------------
# utils  [TS v1.4.0 utils/utils.py, utils/quantize.py, utils/collate.py]
# --------------------------------------------------------------------------------------
def make_ntuple(x, ndim: int) -> Tuple[int, ...]:
    if isinstance(x, int):
        x = tuple(repeat(x, ndim))
    elif isinstance(x, list):
        x = tuple(x)
    assert isinstance(x, tuple) and len(x) == ndim, x
    return x


def ravel_hash(x: np.ndarray) -> np.ndarray:
    assert x.ndim == 2, x.shape
    x = x - np.min(x, axis=0)
    x = x.astype(np.uint64, copy=False)
    xmax = np.max(x, axis=0).astype(np.uint64) + 1
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h += x[:, k]
        h *= xmax[k + 1]
    h += x[:, -1]
    return h


def sparse_quantize(coords, voxel_size=1, *, return_index: bool = False, return_inverse: bool = False):
    """SURVEY A.13; used at core/datasets/semantic_nusc.py:326."""
    if isinstance(voxel_size, (float, int)):
        voxel_size = tuple(repeat(voxel_size, 3))
    voxel_size = np.array(voxel_size)
    coords = np.floor(coords / voxel_size).astype(np.int32)
    _, indices, inverse_indices = np.unique(ravel_hash(coords), return_index=True, return_inverse=True)
    coords = coords[indices]
    outputs = [coords]
    if return_index:
        outputs += [indices]
    if return_inverse:
        outputs += [inverse_indices]
    return outputs[0] if len(outputs) == 1 else outputs


# --------------------------------------------------------------------------------------
# tensors  [TS v1.4.0 tensor.py]   (core/models/utils.py:28-33,59-61,100-108)
# --------------------------------------------------------------------------------------
class SparseTensor:
    def __init__(self, feats, coords, stride=1):
        self.feats = feats
        self.coords = coords
        self.stride = make_ntuple(stride, ndim=3)
        self.cmaps = {}
        self.kmaps = {}

    F = property(lambda s: s.feats, lambda s, v: setattr(s, "feats", v))
    C = property(lambda s: s.coords, lambda s, v: setattr(s, "coords", v))
    s = property(lambda s: s.stride, lambda s, v: setattr(s, "stride", v))

    def cpu(self):
        self.coords = self.coords.cpu()
        self.feats = self.feats.cpu()
        return self

    def cuda(self):
        self.coords = self.coords.cuda()
        self.feats = self.feats.cuda()
        return self

    def detach(self):
        self.coords = self.coords.detach()
        self.feats = self.feats.detach()
        return self

    def to(self, device, non_blocking=True):
        self.coords = self.coords.to(device, non_blocking=non_blocking)
        self.feats = self.feats.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        out = SparseTensor(coords=self.coords, feats=self.feats + other.feats, stride=self.stride)
        out.cmaps = self.cmaps
        out.kmaps = self.kmaps
        return out


class PointTensor:
    def __init__(self, feats, coords, idx_query=None, weights=None):
        self.F = feats
        self.C = coords
        self.idx_query = idx_query if idx_query is not None else {}
        self.weights = weights if weights is not None else {}
        self.additional_features = {"idx_query": {}, "counts": {}}

    def cuda(self):
        self.F = self.F.cuda()
        self.C = self.C.cuda()
        return self

    def detach(self):
        self.F = self.F.detach()
        self.C = self.C.detach()
        return self

    def to(self, device, non_blocking=True):
        self.F = self.F.to(device, non_blocking=non_blocking)
        self.C = self.C.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        out = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        out.additional_features = self.additional_features
        return out


def cat(inputs: List[SparseTensor]) -> SparseTensor:
    """[TS operators.py]; core/models/semantickitti/spvcnn.py:116."""
    feats = torch.cat([x.feats for x in inputs], dim=1)
    out = SparseTensor(coords=inputs[0].coords, feats=feats, stride=inputs[0].stride)
    out.cmaps = inputs[0].cmaps
    out.kmaps = inputs[0].kmaps
    return out


def sparse_collate(inputs: List[SparseTensor]) -> SparseTensor:
    coords, feats = [], []
    stride = inputs[0].stride
    for k, x in enumerate(inputs):
        c, f = x.coords, x.feats
        if isinstance(c, np.ndarray):
            c = torch.tensor(c)
        if isinstance(f, np.ndarray):
            f = torch.tensor(f)
        assert isinstance(c, torch.Tensor) and isinstance(f, torch.Tensor)
        assert x.stride == stride
        b = torch.full((c.shape[0], 1), k, device=c.device, dtype=torch.int)
        coords.append(torch.cat((c, b), dim=1))
        feats.append(f)
    return SparseTensor(coords=torch.cat(coords, dim=0), feats=torch.cat(feats, dim=0), stride=stride)


def sparse_collate_fn(inputs: List) -> dict:
    if isinstance(inputs[0], dict):
        out = {}
        for name in inputs[0].keys():
            v0 = inputs[0][name]
            vals = [x[name] for x in inputs]
            if isinstance(v0, dict):
                out[name] = sparse_collate_fn(vals)
            elif isinstance(v0, np.ndarray):
                out[name] = torch.stack([torch.tensor(v) for v in vals], dim=0)
            elif isinstance(v0, torch.Tensor):
                out[name] = torch.stack(vals, dim=0)
            elif isinstance(v0, SparseTensor):
                out[name] = sparse_collate(vals)
            else:
                out[name] = vals
        return out
    return inputs


# --------------------------------------------------------------------------------------
# nn.utils  [TS v1.4.0 nn/utils/kernel.py, nn/utils/apply.py]
# --------------------------------------------------------------------------------------
def get_kernel_offsets(size, stride=1, dilation=1, device="cpu") -> torch.Tensor:
    """SURVEY A.4; core/models/utils.py:84.  Ordering == weight index."""
    size = make_ntuple(size, ndim=3)
    stride = make_ntuple(stride, ndim=3)
    dilation = make_ntuple(dilation, ndim=3)
    offsets = [np.arange(-size[k] // 2 + 1, size[k] // 2 + 1) * stride[k] * dilation[k] for k in range(3)]
    if np.prod(size) % 2 == 1:
        offsets = [[x, y, z] for z in offsets[2] for y in offsets[1] for x in offsets[0]]
    else:
        offsets = [[x, y, z] for x in offsets[0] for y in offsets[1] for z in offsets[2]]
    return torch.tensor(np.array(offsets), dtype=torch.int, device=device)


def fapply(input: SparseTensor, fn, *args, **kwargs) -> SparseTensor:
    feats = fn(input.feats, *args, **kwargs)
    out = SparseTensor(coords=input.coords, feats=feats, stride=input.stride)
    out.cmaps = input.cmaps
    out.kmaps = input.kmaps
    return out


# --------------------------------------------------------------------------------------
# nn.functional  [TS v1.4.0 nn/functional/*.py over backend/**_cpu.cpp]
# --------------------------------------------------------------------------------------
def sphash(coords: torch.Tensor, offsets: torch.Tensor = None) -> torch.Tensor:
    """SURVEY A.5; core/models/utils.py:19,43,49,86,92."""
    assert coords.dtype == torch.int, coords.dtype
    assert coords.ndim == 2 and coords.shape[1] == 4, coords.shape
    coords = coords.contiguous()
    n = coords.shape[0]
    if offsets is None:
        out = torch.empty(n, dtype=torch.int64)
        _lib().u2o_hash(_p(coords), _i64(n), _p(out))
        return out
    assert offsets.dtype == torch.int, offsets.dtype
    assert offsets.ndim == 2 and offsets.shape[1] == 3, offsets.shape
    offsets = offsets.contiguous()
    K = offsets.shape[0]
    out = torch.empty((K, n), dtype=torch.int64)
    _lib().u2o_kernel_hash(_p(coords), _i64(n), _p(offsets), ctypes.c_int(K), _p(out))
    return out


def sphashquery(queries: torch.Tensor, references: torch.Tensor) -> torch.Tensor:
    """SURVEY A.6; core/models/utils.py:21,50,93,135.  -1 == not found."""
    assert queries.dtype == torch.long and references.dtype == torch.long
    sizes = queries.size()
    q = queries.contiguous().view(-1)
    references = references.contiguous()
    indices = torch.arange(len(references), dtype=torch.long)
    out = torch.empty(q.shape[0], dtype=torch.int64)
    _lib().u2o_hash_query(_p(q), _i64(q.shape[0]), _p(references), _p(indices), _i64(len(references)), _p(out))
    return (out - 1).view(*sizes)


def spcount(coords: torch.Tensor, num: int) -> torch.Tensor:
    """SURVEY A.7; core/models/utils.py:22,51."""
    assert coords.dtype == torch.int
    coords = coords.contiguous()
    out = torch.empty(int(num), dtype=torch.int)
    _lib().u2o_count(_p(coords), _i64(coords.shape[0]), _p(out), _i64(num))
    return out


class _VoxelizeFn(Function):
    """SURVEY A.8 [TS nn/functional/voxelize.py]; core/models/utils.py:24,26,58."""

    @staticmethod
    def forward(ctx, feats, coords, counts):
        feats = feats.contiguous()
        coords = coords.contiguous().int()
        counts = counts.contiguous().int()
        N, c = feats.shape
        s = counts.shape[0]
        out = torch.empty((s, c), dtype=feats.dtype)
        getattr(_lib(), "u2o_voxelize_fwd_" + _suf(feats))(_p(feats), _i64(N), _i64(c), _p(coords), _p(counts), _p(out), _i64(s))
        ctx.for_backwards = (coords, counts, N)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        coords, counts, N = ctx.for_backwards
        grad_output = grad_output.contiguous()
        s, c = grad_output.shape
        gin = torch.empty((N, c), dtype=grad_output.dtype)
        getattr(_lib(), "u2o_voxelize_bwd_" + _suf(grad_output))(_p(grad_output), _i64(N), _i64(c), _p(coords), _p(counts), _p(gin), _i64(s))
        return gin, None, None


def spvoxelize(feats, coords, counts):
    return _VoxelizeFn.apply(feats, coords, counts)


def calc_ti_weights(coords: torch.Tensor, idx_query: torch.Tensor, scale: float = 1) -> torch.Tensor:
    """SURVEY A.9 [TS nn/functional/devoxelize.py]; core/models/utils.py:94."""
    with torch.no_grad():
        p = coords
        if scale != 1:
            pf = torch.floor(coords / scale) * scale
        else:
            pf = torch.floor(coords)
        pc = pf + scale
        x, y, z = p[:, 0].view(-1, 1), p[:, 1].view(-1, 1), p[:, 2].view(-1, 1)
        xf, yf, zf = pf[:, 0].view(-1, 1).float(), pf[:, 1].view(-1, 1).float(), pf[:, 2].view(-1, 1).float()
        xc, yc, zc = pc[:, 0].view(-1, 1).float(), pc[:, 1].view(-1, 1).float(), pc[:, 2].view(-1, 1).float()
        w0 = (xc - x) * (yc - y) * (zc - z)
        w1 = (xc - x) * (yc - y) * (z - zf)
        w2 = (xc - x) * (y - yf) * (zc - z)
        w3 = (xc - x) * (y - yf) * (z - zf)
        w4 = (x - xf) * (yc - y) * (zc - z)
        w5 = (x - xf) * (yc - y) * (z - zf)
        w6 = (x - xf) * (y - yf) * (zc - z)
        w7 = (x - xf) * (y - yf) * (z - zf)
        w = torch.cat([w0, w1, w2, w3, w4, w5, w6, w7], dim=1)
        w = w.transpose(1, 0).contiguous()
        if scale != 1:
            w /= scale ** 3
        w[idx_query == -1] = 0
        w /= torch.sum(w, dim=0) + 1e-8
    return w


class _DevoxelizeFn(Function):
    """SURVEY A.9; core/models/utils.py:99,111."""

    @staticmethod
    def forward(ctx, feats, coords, weights):
        feats = feats.contiguous()
        coords = coords.contiguous().int()
        weights = weights.contiguous().to(feats.dtype)
        n, c = feats.shape
        N = coords.shape[0]
        out = torch.empty((N, c), dtype=feats.dtype)
        getattr(_lib(), "u2o_devoxelize_fwd_" + _suf(feats))(_p(feats), _i64(n), _i64(c), _p(coords), _p(weights), _i64(N), _p(out))
        ctx.for_backwards = (coords, weights, n)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        coords, weights, n = ctx.for_backwards
        grad_output = grad_output.contiguous()
        N, c = grad_output.shape
        g = torch.empty((n, c), dtype=grad_output.dtype)
        getattr(_lib(), "u2o_devoxelize_bwd_" + _suf(grad_output))(_p(grad_output), _i64(N), _i64(c), _p(coords), _p(weights), _i64(n), _p(g))
        return g, None, None


def spdevoxelize(feats, coords, weights):
    return _DevoxelizeFn.apply(feats, coords, weights)


def spdownsample(coords: torch.Tensor, stride=2, kernel_size=2, tensor_stride=1) -> torch.Tensor:
    """SURVEY A.10 [TS nn/functional/downsample.py]; output sorted by (b,x,y,z)."""
    stride = make_ntuple(stride, ndim=3)
    kernel_size = make_ntuple(kernel_size, ndim=3)
    tensor_stride = make_ntuple(tensor_stride, ndim=3)
    sample_stride = torch.tensor([stride[k] * tensor_stride[k] for k in range(3)], dtype=torch.int).unsqueeze(0)
    if all(stride[k] in [1, kernel_size[k]] for k in range(3)):
        coords = coords.clone()
        coords[:, :3] = torch.div(coords[:, :3], sample_stride, rounding_mode="floor") * sample_stride
    else:
        offsets = get_kernel_offsets(kernel_size, tensor_stride)
        kv = offsets.size(0)
        cmin = torch.min(coords[:, :3], dim=0, keepdim=True).values
        x = coords[:, :3].unsqueeze(1).repeat(1, kv, 1) + offsets
        b = coords[:, 3:].repeat(1, kv)
        coords = torch.cat([x.view(-1, 3), b.view(-1, 1)], dim=1)
        mask = (coords[:, :3] % sample_stride == 0)
        mask &= (coords[:, :3] >= cmin)
        coords = coords[torch.all(mask, dim=1)]
    coords = coords[:, [3, 0, 1, 2]]
    coords = torch.unique(coords, dim=0)
    coords = coords[:, [1, 2, 3, 0]]
    return coords


# Yardstick for the reduced-precision modes of the product (tests/gradtable.py): with OPERAND_ROUNDING set, every conv
# GEMM operand (features, weights, output gradients) is rounded to that format before the mm — the same rounding points as
# the product's bf16 / tf32 tensor-core modes — while everything else keeps the oracle's dtype.  None = the plain oracle.
OPERAND_ROUNDING = None


def _rnd(x: torch.Tensor) -> torch.Tensor:
    if OPERAND_ROUNDING is None:
        return x
    if OPERAND_ROUNDING == "bf16":
        return x.bfloat16().to(x.dtype)
    if OPERAND_ROUNDING == "tf32":  # the tensor core truncates fp32 to 10 mantissa bits
        return (x.float().contiguous().view(torch.int32) & -8192).view(torch.float32).to(x.dtype)
    raise ValueError(OPERAND_ROUNDING)


class _ConvolutionFn(Function):
    """SURVEY A.11/A.12 [TS nn/functional/conv.py, backend/convolution/convolution_cpu.cpp]:
    per kernel offset gather -> mm -> scatter; centre-tap shortcut iff K odd and N_in == N_out."""

    @staticmethod
    def forward(ctx, input, weight, nbmaps, nbsizes, sizes, transposed=False):
        input = _rnd(input.contiguous())
        weight = _rnd(weight.contiguous())
        nbmaps = nbmaps.int().contiguous()
        nbsizes = nbsizes.int().contiguous()
        n_out = sizes[1] if not transposed else sizes[0]
        output = torch.zeros(n_out, weight.size(-1), dtype=input.dtype)
        suf = _suf(input)
        t = int(bool(transposed))
        K = weight.shape[0]
        cin, cout = weight.shape[1], weight.shape[2]
        mid = K // 2
        pre_mid = (K % 2 == 1) and (input.shape[0] == output.shape[0])
        if pre_mid:
            torch.mm(input, weight[mid], out=output)
        cur = 0
        sizes_list = nbsizes.tolist()
        for k in range(K):
            na = sizes_list[k]
            if na == 0:
                continue
            if k == mid and pre_mid:
                cur += na
                continue
            nb = nbmaps[cur:cur + na]
            buf = torch.empty((na, cin), dtype=input.dtype)
            getattr(_lib(), "u2o_gather_" + suf)(_p(input), _i64(cin), _p(nb), _i64(na), ctypes.c_int(t), _p(buf))
            prod = torch.mm(buf, weight[k])
            getattr(_lib(), "u2o_scatter_" + suf)(_p(prod), _i64(cout), _p(nb), _i64(na), ctypes.c_int(t), _p(output))
            cur += na
        ctx.for_backwards = (input, weight, nbmaps, sizes_list, t)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, nbmaps, sizes_list, t = ctx.for_backwards
        grad_output = _rnd(grad_output.contiguous())
        suf = _suf(input)
        K, cin, cout = weight.shape
        grad_input = torch.zeros_like(input)
        grad_weight = torch.zeros_like(weight)
        cur = 0
        for k in range(K):
            na = sizes_list[k]
            if na == 0:
                continue
            nb = nbmaps[cur:cur + na]
            gbuf = torch.empty((na, cout), dtype=input.dtype)
            ibuf = torch.empty((na, cin), dtype=input.dtype)
            # gather grad_out with flag !t, input with flag t
            getattr(_lib(), "u2o_gather_" + suf)(_p(grad_output), _i64(cout), _p(nb), _i64(na), ctypes.c_int(1 - t), _p(gbuf))
            getattr(_lib(), "u2o_gather_" + suf)(_p(input), _i64(cin), _p(nb), _i64(na), ctypes.c_int(t), _p(ibuf))
            gi = torch.mm(gbuf, weight[k].t())
            torch.mm(ibuf.t(), gbuf, out=grad_weight[k])
            getattr(_lib(), "u2o_scatter_" + suf)(_p(gi.contiguous()), _i64(cin), _p(nb), _i64(na), ctypes.c_int(1 - t), _p(grad_input))
            cur += na
        return grad_input, grad_weight, None, None, None, None


def build_kernel_map(coords, in_stride, kernel_size, stride, dilation):
    """Kernel-map construction inside F.conv3d.  Returns
    (kmap=[nbmaps int64 [M,2] (in,out) sorted by (k,out), nbsizes [K], (N_in,N_out)], out_coords)."""
    offsets = get_kernel_offsets(kernel_size, stride=in_stride)
    references = sphash(coords)
    out_coords = coords
    if any(s > 1 for s in stride):
        out_coords = spdownsample(coords, stride, kernel_size, in_stride)
    queries = sphash(out_coords, offsets)
    results = sphashquery(queries, references)
    nbsizes = torch.sum(results != -1, dim=1)
    nbmaps = torch.nonzero(results != -1)
    indices = nbmaps[:, 0] * results.size(1) + nbmaps[:, 1]
    nbmaps[:, 0] = results.view(-1)[indices]
    return [nbmaps, nbsizes, (coords.shape[0], out_coords.shape[0])], out_coords


def conv3d(input: SparseTensor, weight, kernel_size, bias=None, stride=1, dilation=1, transposed=False) -> SparseTensor:
    feats, coords = input.feats, input.coords
    kernel_size = make_ntuple(kernel_size, ndim=3)
    stride = make_ntuple(stride, ndim=3)
    dilation = make_ntuple(dilation, ndim=3)
    if kernel_size == (1, 1, 1) and stride == (1, 1, 1) and dilation == (1, 1, 1):
        feats = feats.matmul(weight)
        if bias is not None:
            feats = feats + bias
        output = SparseTensor(coords=coords, feats=feats, stride=input.stride)
    elif not transposed:
        key = (input.stride, kernel_size, stride, dilation)
        kmap = input.kmaps.get(key)
        if kmap is None:
            kmap, coords = build_kernel_map(coords, input.stride, kernel_size, stride, dilation)
            input.kmaps[key] = kmap
        elif any(s > 1 for s in stride):
            coords = input.cmaps[tuple(input.stride[k] * stride[k] for k in range(3))]
        feats = _ConvolutionFn.apply(feats, weight, kmap[0], kmap[1], kmap[2], transposed)
        if bias is not None:
            feats = feats + bias
        output = SparseTensor(coords=coords, feats=feats, stride=tuple(input.stride[k] * stride[k] for k in range(3)))
    else:
        tensor_stride = tuple(input.stride[k] // stride[k] for k in range(3))
        kmap = input.kmaps[(tensor_stride, kernel_size, stride, dilation)]
        feats = _ConvolutionFn.apply(feats, weight, kmap[0], kmap[1], kmap[2], transposed)
        if bias is not None:
            feats = feats + bias
        output = SparseTensor(coords=input.cmaps[tensor_stride], feats=feats, stride=tensor_stride)
    output.cmaps = input.cmaps
    output.cmaps.setdefault(output.stride, output.coords)
    output.kmaps = input.kmaps
    return output


# --------------------------------------------------------------------------------------
# nn modules  [TS v1.4.0 nn/modules/{conv,norm,activation}.py]
# --------------------------------------------------------------------------------------
class Conv3d(nn.Module):
    """SURVEY A.3; ctor sites core/models/build_blocks.py:25-29,43-47,59-70,76."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False, transposed=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = make_ntuple(kernel_size, ndim=3)
        self.stride = make_ntuple(stride, ndim=3)
        self.dilation = dilation
        self.transposed = transposed
        self.kernel_volume = int(np.prod(self.kernel_size))
        if self.kernel_volume > 1:
            self.kernel = nn.Parameter(torch.zeros(self.kernel_volume, in_channels, out_channels))
        else:
            self.kernel = nn.Parameter(torch.zeros(in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        import math
        std = 1 / math.sqrt((self.out_channels if self.transposed else self.in_channels) * self.kernel_volume)
        self.kernel.data.uniform_(-std, std)
        if self.bias is not None:
            self.bias.data.uniform_(-std, std)

    def forward(self, input):
        return conv3d(input, self.kernel, kernel_size=self.kernel_size, bias=self.bias, stride=self.stride,
                      dilation=self.dilation, transposed=self.transposed)


class BatchNorm(nn.BatchNorm1d):
    def forward(self, input):
        return fapply(input, super().forward)


class ReLU(nn.ReLU):
    def forward(self, input):
        return fapply(input, super().forward)


# --------------------------------------------------------------------------------------
# namespace objects mirroring the torchsparse module tree, so that model code written
# against `torchsparse` can be pointed at the oracle (tests only).
# --------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def as_torchsparse_modules(prefix: str = "torchsparse"):
    """Build module objects {name: module} exposing this oracle under torchsparse's layout."""
    functional = _mod(prefix + ".nn.functional", sphash=sphash, sphashquery=sphashquery, spcount=spcount,
                      spvoxelize=spvoxelize, spdevoxelize=spdevoxelize, calc_ti_weights=calc_ti_weights,
                      spdownsample=spdownsample, conv3d=conv3d)
    nn_utils = _mod(prefix + ".nn.utils", get_kernel_offsets=get_kernel_offsets, fapply=fapply)
    nn_mod = _mod(prefix + ".nn", Conv3d=Conv3d, BatchNorm=BatchNorm, ReLU=ReLU, functional=functional, utils=nn_utils)
    quantize = _mod(prefix + ".utils.quantize", sparse_quantize=sparse_quantize, ravel_hash=ravel_hash)
    collate = _mod(prefix + ".utils.collate", sparse_collate=sparse_collate, sparse_collate_fn=sparse_collate_fn)
    utils = _mod(prefix + ".utils", make_ntuple=make_ntuple, quantize=quantize, collate=collate)
    top = _mod(prefix, SparseTensor=SparseTensor, PointTensor=PointTensor, cat=cat, nn=nn_mod, utils=utils)
    top.__path__ = []
    nn_mod.__path__ = []
    utils.__path__ = []
    return {prefix: top, prefix + ".nn": nn_mod, prefix + ".nn.functional": functional,
            prefix + ".nn.utils": nn_utils, prefix + ".utils": utils, prefix + ".utils.quantize": quantize,
            prefix + ".utils.collate": collate}


def install_as_torchsparse():
    """TESTS ONLY: register the oracle as `torchsparse` in sys.modules."""
    mods = as_torchsparse_modules()
    sys.modules.update(mods)
    return mods["torchsparse"]
