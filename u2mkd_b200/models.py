"""Host-side mirror of the reference's point-voxel glue and SPVCNN backbone, written against
a torchsparse-like namespace so the same code drives the CUDA path (product) and, in tests,
the CPU oracle.

Mirrors (same names, argument meaning, module tree and parameter names, so state_dicts are
interchangeable with the reference's):
  initial_voxelize / point_to_voxel / voxel_to_point / fetch_idx  <- core/models/utils.py:15-135
  BasicConvolutionBlock / BasicDeconvolutionBlock / ResidualBlock  <- core/models/build_blocks.py:21-84
  SPVCNN                                                           <- core/models/semantickitti/spvcnn.py:10-142
The reference files themselves also run unchanged on u2mkd_b200.torchsparse
(`u2mkd_b200.install_as_torchsparse()`); this module exists because /root/reference is not
present on the GPU box where bench.py and the -m gpu tests run.
"""
from __future__ import annotations

import weakref
from types import SimpleNamespace

import torch
from torch import nn

CHANNELS = (32, 32, 64, 128, 256, 256, 128, 96, 96)


class PreparedScan:
    """What prepare_scan_begin / _finish know about a scan batch: x (the input SparseTensor), res, and after phase B
    idx_query / counts / coords / cmaps / kmaps / done (event on the prefetch stream)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def build_family(ts) -> SimpleNamespace:
    """Instantiate the glue functions and model classes over `ts` (a torchsparse-like module)."""
    spnn, spf = ts.nn, ts.nn.functional
    SparseTensor, PointTensor = ts.SparseTensor, ts.PointTensor
    get_kernel_offsets = ts.nn.utils.get_kernel_offsets
    coord_query = getattr(spf, "coord_query", None)  # product-only fast path (cached per-stride coordinate table)

    # ------------------------------------------------------------------ point <-> voxel
    def _floor_to_stride(z, s):
        """Voxel key of every point at stride s: (floor(xyz / s) * s, batch), int32 [N,4]."""
        return torch.cat([torch.floor(z.C[:, :3] / s).int() * s, z.C[:, -1].int().view(-1, 1)], 1)

    def _voxel_keys(zc, init_res, after_res):
        new_float_coord = torch.cat([(zc[:, :3] * init_res) / after_res, zc[:, -1].view(-1, 1)], 1)
        return new_float_coord, torch.floor(new_float_coord)

    def prepare_scan_begin(x_in, init_res, after_res):
        """Product path only, phase A of the coordinate prefetch for the scan batch `x_in` (the SparseTensor that will be
        passed as in_mod["lidar"], or a callable returning it — e.g. the host -> device copies): call it BEFORE queueing the
        current step.  Launches voxel keys, unique and the coarse coordinate sets on the prefetch stream; no host wait."""
        from . import ops

        def work():
            x = x_in() if callable(x_in) else x_in
            _nfc, floored = _voxel_keys(x.C.float(), init_res, after_res)
            return PreparedScan(x=x, src=x.C, res=(init_res, after_res), a=ops.coords_begin(floored.int()), done=None)

        return ops.coord_prefetch.begin(work)

    def prepare_scan_finish(prep):
        """Phase B: call it AFTER queueing the current step.  Reads the row counts phase A left in pinned memory and queues
        every kernel map / tile sort / pair list an earlier forward asked for.  prep.x is the batch to pass to the model (keep
        `prep` alive until that forward pass has run); initial_voxelize picks the results up (same values as in place)."""
        from . import ops

        def work():
            prep.idx_query, prep.counts, prep.coords, prep.cmaps, prep.kmaps = ops.coords_finish(prep.a, SparseTensor)
            prep.a = None
            return prep

        _, prep.done = ops.coord_prefetch.finish(work)
        # weak: prep holds x; a strong reference back would make a cycle that only the cyclic collector frees — device memory
        # of old batches would be released late and at random (measured: the prefetch pool grew by 370 MB per step)
        prep.x._u2_prep = weakref.ref(prep)
        return prep

    def prepare_scan(x_in, init_res, after_res):
        """Both phases back to back (blocks until phase A's kernels have run)."""
        return prepare_scan_finish(prepare_scan_begin(x_in, init_res, after_res))

    def initial_voxelize(z, init_res, after_res):
        """utils.py:15-35 — quantise, hash, unique, scatter-mean coords + features."""
        new_float_coord, floored = _voxel_keys(z.C, init_res, after_res)
        fused = getattr(spf, "unique_voxelize", None)
        prep = getattr(z, "_u2_prep", None)
        cmaps = kmaps = None
        if prep is not None and prep.done is not None and prep.res == (init_res, after_res) and floored.is_cuda:
            # product path with prepare_scan_*(): the index part and the kernel maps were computed on the prefetch stream
            torch.cuda.current_stream().wait_event(prep.done)
            idx_query, counts, coords, cmaps, kmaps = prep.idx_query, prep.counts, prep.coords, prep.cmaps, prep.kmaps
        elif fused is not None and floored.is_cuda:
            # product path: the five index operators below as one call (same voxel order, same idx_query / counts / coords),
            # the coarse coordinate sets next to it and ONE host wait for all row counts, then every planned kernel map
            from . import ops
            idx_query, counts, coords, cmaps, kmaps = ops.coords_finish(ops.coords_begin(floored.int()), SparseTensor)
        else:
            pc_hash = spf.sphash(floored.int())
            sparse_hash = torch.unique(pc_hash)
            idx_query = spf.sphashquery(pc_hash, sparse_hash)
            counts = spf.spcount(idx_query.int(), len(sparse_hash))
            coords = torch.round(spf.spvoxelize(floored, idx_query, counts)).int()
        feats = spf.spvoxelize(z.F, idx_query, counts)
        x = SparseTensor(feats, coords, 1)
        if cmaps is not None:
            x.cmaps, x.kmaps = cmaps, kmaps
        x.cmaps.setdefault(x.stride, x.coords)
        prebuild = getattr(spf, "prebuild_maps", None)
        if prebuild is not None and coords.is_cuda and cmaps is None:
            prebuild(x)  # (a torchsparse namespace with prebuild_maps but without unique_voxelize)
        z.additional_features["idx_query"][1] = idx_query
        z.additional_features["counts"][1] = counts
        z.C = new_float_coord
        return x

    def point_to_voxel(x, z):
        """utils.py:40-65 — scatter-mean point features into the voxels of x (cached per stride)."""
        cache = z.additional_features
        if cache is None or cache.get("idx_query") is None or cache["idx_query"].get(x.s) is None:
            if coord_query is not None and x.C.is_cuda:  # product path: one lookup kernel on the stride's cached table
                idx_query = coord_query(_floor_to_stride(z, x.s[0]), x.C)
            else:
                idx_query = spf.sphashquery(spf.sphash(_floor_to_stride(z, x.s[0])), spf.sphash(x.C))
            counts = spf.spcount(idx_query.int(), x.C.shape[0])
            cache["idx_query"][x.s] = idx_query
            cache["counts"][x.s] = counts
        else:
            idx_query, counts = cache["idx_query"][x.s], cache["counts"][x.s]
        out = SparseTensor(spf.spvoxelize(z.F, idx_query, counts), x.C, x.s)
        out.cmaps = x.cmaps
        out.kmaps = x.kmaps
        return out

    def voxel_to_point(x, z, nearest=False):
        """utils.py:70-118 — trilinear devoxelise of x onto the points of z (cached per stride)."""
        if z.idx_query is None or z.weights is None or z.idx_query.get(x.s) is None or z.weights.get(x.s) is None:
            off = get_kernel_offsets(2, x.s, 1, device=z.F.device)
            if coord_query is not None and x.C.is_cuda and z.F.is_cuda:
                idx_query = coord_query(_floor_to_stride(z, x.s[0]), x.C, off)
            else:
                old_hash = spf.sphash(_floor_to_stride(z, x.s[0]), off)
                idx_query = spf.sphashquery(old_hash, spf.sphash(x.C.to(z.F.device)))
            weights = spf.calc_ti_weights(z.C, idx_query, scale=x.s[0]).transpose(0, 1).contiguous()
            idx_query = idx_query.transpose(0, 1).contiguous()
            if nearest:
                weights[:, 1:] = 0.
                idx_query[:, 1:] = -1
            new_feat = spf.spdevoxelize(x.F, idx_query, weights)
            out = PointTensor(new_feat, z.C, idx_query=z.idx_query, weights=z.weights)
            out.additional_features = z.additional_features
            out.idx_query[x.s] = idx_query
            out.weights[x.s] = weights
            z.idx_query[x.s] = idx_query
            z.weights[x.s] = weights
        else:
            new_feat = spf.spdevoxelize(x.F, z.idx_query.get(x.s), z.weights.get(x.s))
            out = PointTensor(new_feat, z.C, idx_query=z.idx_query, weights=z.weights)
            out.additional_features = z.additional_features
        return out

    def fetch_idx(source_coords, target_coords):
        """utils.py:121-135 — row of every source coordinate in target_coords, -1 if absent."""
        assert isinstance(source_coords, torch.Tensor) and isinstance(target_coords, torch.Tensor)
        return spf.sphashquery(spf.sphash(source_coords), spf.sphash(target_coords))

    class SparseSyncBatchNorm(nn.SyncBatchNorm):
        """utils.py:138-220 — SyncBatchNorm over SparseTensor.F + recursive module conversion."""

        def forward(self, input):
            return ts.nn.utils.fapply(input, super().forward)

        @classmethod
        def convert_sync_batchnorm(cls, module, process_group=None):
            converted = module
            target = cls if isinstance(module, spnn.BatchNorm) else (
                nn.SyncBatchNorm if isinstance(module, nn.modules.batchnorm._BatchNorm) else None)
            if target is not None:
                converted = target(module.num_features, module.eps, module.momentum, module.affine,
                                   module.track_running_stats, process_group)
                if module.affine:
                    with torch.no_grad():
                        converted.weight = module.weight
                        converted.bias = module.bias
                converted.running_mean = module.running_mean
                converted.running_var = module.running_var
                converted.num_batches_tracked = module.num_batches_tracked
            for name, child in module.named_children():
                converted.add_module(name, cls.convert_sync_batchnorm(child, process_group))
            return converted

    # ------------------------------------------------------------------ blocks
    def _conv_bn(inc, outc, ks, stride=1, dilation=1, transposed=False, relu=True):
        layers = [spnn.Conv3d(inc, outc, kernel_size=ks, stride=stride, dilation=dilation, transposed=transposed),
                  spnn.BatchNorm(outc)]
        if relu:
            layers.append(spnn.ReLU(True))
        return layers

    class BasicConvolutionBlock(nn.Module):
        def __init__(self, inc, outc, ks=3, stride=1, dilation=1):
            super().__init__()
            self.net = nn.Sequential(*_conv_bn(inc, outc, ks, stride, dilation))

        def forward(self, x):
            return self.net(x)

    class BasicDeconvolutionBlock(nn.Module):
        def __init__(self, inc, outc, ks=3, stride=1):
            super().__init__()
            self.net = nn.Sequential(*_conv_bn(inc, outc, ks, stride, transposed=True))

        def forward(self, x):
            return self.net(x)

    class ResidualBlock(nn.Module):
        def __init__(self, inc, outc, ks=3, stride=1, dilation=1):
            super().__init__()
            self.net = nn.Sequential(*(_conv_bn(inc, outc, ks, stride, dilation) +
                                       _conv_bn(outc, outc, ks, 1, dilation, relu=False)))
            if inc == outc and stride == 1:
                self.downsample = nn.Sequential()
            else:
                self.downsample = nn.Sequential(*_conv_bn(inc, outc, 1, stride, 1, relu=False))
            self.relu = spnn.ReLU(True)

        def forward(self, x):
            return self.relu(self.net(x) + self.downsample(x))

    # ------------------------------------------------------------------ SPVCNN
    class SPVCNN(nn.Module):
        """Point-voxel U-Net: stem, 4 strided stages down, 4 transposed stages up, 3 point MLPs."""

        def __init__(self, **kwargs):
            super().__init__()
            cr = kwargs.get("cr", 1.0)
            cs = [int(cr * c) for c in CHANNELS]
            self.in_channel = kwargs.get("in_channel", 4)
            self.num_classes = kwargs.get("num_classes", 17)
            self.out_channel = cs[-1]
            if "pres" in kwargs and "vres" in kwargs:
                self.pres, self.vres = kwargs["pres"], kwargs["vres"]

            self.stem = nn.Sequential(*(_conv_bn(self.in_channel, cs[0], 3) + _conv_bn(cs[0], cs[0], 3)))
            self.vox_downs = nn.ModuleList(
                nn.Sequential(BasicConvolutionBlock(cs[i], cs[i], ks=2, stride=2, dilation=1),
                              ResidualBlock(cs[i], cs[i + 1], ks=3, stride=1, dilation=1),
                              ResidualBlock(cs[i + 1], cs[i + 1], ks=3, stride=1, dilation=1))
                for i in range(4))
            self.vox_ups = nn.ModuleList(
                nn.ModuleList([
                    BasicDeconvolutionBlock(cs[i], cs[i + 1], ks=2, stride=2),
                    nn.Sequential(ResidualBlock(cs[i + 1] + cs[7 - i], cs[i + 1], ks=3, stride=1, dilation=1),
                                  ResidualBlock(cs[i + 1], cs[i + 1], ks=3, stride=1, dilation=1))])
                for i in range(4, 8))
            self.classifier_vox = nn.Sequential(nn.Linear(cs[8], self.num_classes))
            self.point_transforms = nn.ModuleList(
                nn.Sequential(nn.Linear(a, b), nn.BatchNorm1d(b), nn.ReLU(True))
                for a, b in ((cs[0], cs[4]), (cs[4], cs[6]), (cs[6], cs[8])))
            for m in self.modules():
                if isinstance(m, nn.BatchNorm1d):
                    nn.init.constant_(m.weight, 1)
                    nn.init.constant_(m.bias, 0)
            self.dropout = nn.Dropout(0.3, True)

        def _up(self, stage, y, skip):
            y = self.vox_ups[stage][0](y)
            return self.vox_ups[stage][1](ts.cat([y, skip]))

        def forward(self, in_mod):
            x = in_mod["lidar"]
            z = PointTensor(x.F, x.C.float())
            prep = getattr(x, "_u2_prep", None)
            prep = prep() if prep is not None else None
            if prep is not None and prep.src is x.C:
                z._u2_prep = prep   # prepare_scan() ran for this batch
            x0 = self.stem(initial_voxelize(z, self.pres, self.vres))
            z0 = voxel_to_point(x0, z, nearest=False)

            feats = [point_to_voxel(x0, z0)]
            for down in self.vox_downs:
                feats.append(down(feats[-1]))
            x1, x2, x3, x4 = feats[1:]

            z1 = voxel_to_point(x4, z0)
            z1.F = z1.F + self.point_transforms[0](z0.F)

            y1 = point_to_voxel(x4, z1)
            y1.F = self.dropout(y1.F)
            y1 = self._up(0, y1, x3)
            y2 = self._up(1, y1, x2)
            z2 = voxel_to_point(y2, z1)
            z2.F = z2.F + self.point_transforms[1](z1.F)

            y3 = point_to_voxel(y2, z2)
            y3.F = self.dropout(y3.F)
            y3 = self._up(2, y3, x1)
            y4 = self._up(3, y3, x0)
            z3 = voxel_to_point(y4, z2)
            z3.F = z3.F + self.point_transforms[2](z2.F)
            return {"x_vox": self.classifier_vox(z3.F)}

    return SimpleNamespace(initial_voxelize=initial_voxelize, prepare_scan=prepare_scan, prepare_scan_begin=prepare_scan_begin,
                           prepare_scan_finish=prepare_scan_finish, point_to_voxel=point_to_voxel,
                           voxel_to_point=voxel_to_point, fetch_idx=fetch_idx,
                           SparseSyncBatchNorm=SparseSyncBatchNorm,
                           BasicConvolutionBlock=BasicConvolutionBlock,
                           BasicDeconvolutionBlock=BasicDeconvolutionBlock, ResidualBlock=ResidualBlock,
                           SPVCNN=SPVCNN, ts=ts)


_product = None


def product() -> SimpleNamespace:
    """The family bound to the CUDA path (u2mkd_b200.torchsparse)."""
    global _product
    if _product is None:
        from . import torchsparse as ts
        _product = build_family(ts)
    return _product
