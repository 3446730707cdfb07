"""Point <-> pixel transforms of the student's fusion path on the C-ABI kernels of csrc/pixel.cu, with
the reference's function names and arguments:

    Point2Grid(pts_feat, pixel_coordinates, masks, grid_size)       core/models/fusion_blocks.py:217-238
    Feature_Gather(feature_map, xy, mode='bilinear')                 core/models/fusion_blocks.py:241-254
    Feature_Fetch(masks, pix_coord, imfeats, mode='bilinear')        core/models/fusion_blocks.py:257-278
    multiscale_point2grid(...)    the loop body of core/models/nuscenes/spvcnn_swiftnet18_spformer_tsd_full.py:448-478

The reference loops over batch elements, cameras and scales in Python, each iteration a torch.unique(dim=0) (sort + host
sync), a scatter_add_, a sparse_coo_tensor().to_dense() and a permute; here all cameras of a batch element go through one
scatter kernel + one normalise/transpose kernel per scale, and one gather kernel, with autograd Functions on top."""
import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import ops
from ._lib import check, lib


class _Point2GridFn(Function):
    @staticmethod
    def forward(ctx, feats, coord, mask, H, W):
        ops._need_cuda(feats, coord, mask)
        feats = feats.contiguous().float()
        coord = coord.contiguous().float()
        mask8 = mask.contiguous().to(torch.uint8)
        V, N = mask8.shape
        C = feats.shape[1]
        assert coord.shape == (V, N, 2) and feats.shape[0] == N, (coord.shape, feats.shape, mask8.shape)
        grid = torch.empty((V, C, H, W), dtype=torch.float32, device=feats.device)
        counts = torch.empty((V, H, W), dtype=torch.int32, device=feats.device)
        sbytes = lib().u2_point2grid_scratch_bytes(C, V, H, W)
        scratch = ops._ws("p2g", sbytes, feats.device)
        check(lib().u2_point2grid_fwd(feats.data_ptr(), coord.data_ptr(), mask8.data_ptr(), N, C, V, H, W, grid.data_ptr(),
                                      counts.data_ptr(), scratch.data_ptr(), scratch.numel(), ops._st()))
        ops._count(2)
        ctx.save_for_backward(coord, mask8, counts)
        ctx.shape = (N, C, V, H, W)
        return grid

    @staticmethod
    def backward(ctx, dgrid):
        coord, mask8, counts = ctx.saved_tensors
        N, C, V, H, W = ctx.shape
        dgrid = dgrid.contiguous().float()
        dfeats = torch.empty((N, C), dtype=torch.float32, device=dgrid.device)
        check(lib().u2_point2grid_bwd(dgrid.data_ptr(), coord.data_ptr(), mask8.data_ptr(), counts.data_ptr(), N, C, V, H, W,
                                      dfeats.data_ptr(), ops._st()))
        ops._count()
        return dfeats, None, None, None, None


def point2grid(feats, coord, mask, grid_size):
    """Per-pixel mean of the masked points' features for the V cameras of ONE batch element: feats [N, C], coord [V, N, 2]
    in [-1, 1], mask bool [V, N] -> [V, C, H, W] (zeros where no point falls)."""
    H, W = grid_size
    return _Point2GridFn.apply(feats, coord, mask, int(H), int(W))


def Point2Grid(pts_feat, pixel_coordinates, masks, grid_size):
    """core/models/fusion_blocks.py:217-238: lists over the batch of coord [V, N_b, 2] / mask [V, N_b]; points of the batch
    elements are stacked in pts_feat.  Returns [sum_b V_b, C, H, W]."""
    cur, out = 0, []
    for mask, coord in zip(masks, pixel_coordinates):
        n = mask.size(1)
        out.append(point2grid(pts_feat[cur:cur + n], coord, mask, grid_size))
        cur += n
    return torch.cat(out, 0)


def multiscale_point2grid(pts_feat, pixel_coordinates, masks, grid_size, n_scales):
    """Loop body of spvcnn_swiftnet18_spformer_tsd_full.py:448-478: the scatter-mean at n_scales resolutions
    (H / 2^s, W / 2^s rounded as there), each bilinearly upsampled (align_corners=True, build_blocks.py:18) to grid_size,
    averaged over the scales; a camera that sees no point contributes zeros."""
    ifh, ifw = grid_size
    cur, out = 0, []
    for mask, coord in zip(masks, pixel_coordinates):
        n = mask.size(1)
        feats = pts_feat[cur:cur + n]
        acc, cnt = None, 1
        for _ in range(n_scales):
            c_ih, c_iw = int(round(float(ifh) / cnt + 0.01)), int(round(float(ifw) / cnt + 0.01))
            g = point2grid(feats, coord, mask, (c_ih, c_iw))
            if (c_ih, c_iw) != (ifh, ifw):
                g = F.interpolate(g, (ifh, ifw), mode="bilinear", align_corners=True)
            acc = g if acc is None else acc + g
            cnt *= 2
        out.append(acc / n_scales)
        cur += n
    return torch.cat(out, 0).contiguous()


class _PixelGatherFn(Function):
    @staticmethod
    def forward(ctx, img, coord, mask):
        ops._need_cuda(img, coord, mask)
        img = img.contiguous().float()
        coord = coord.contiguous().float()
        mask8 = mask.contiguous().to(torch.uint8)
        V, C, H, W = img.shape
        N = mask8.shape[1]
        assert coord.shape == (V, N, 2) and mask8.shape[0] == V
        out = torch.empty((N, C), dtype=torch.float32, device=img.device)
        check(lib().u2_pixel_gather_fwd(img.data_ptr(), coord.data_ptr(), mask8.data_ptr(), N, C, V, H, W, out.data_ptr(), ops._st()))
        ops._count()
        ctx.save_for_backward(coord, mask8)
        ctx.shape = (N, C, V, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        coord, mask8 = ctx.saved_tensors
        N, C, V, H, W = ctx.shape
        dout = dout.contiguous().float()
        dimg = torch.empty((V, C, H, W), dtype=torch.float32, device=dout.device)
        check(lib().u2_pixel_gather_bwd(dout.data_ptr(), coord.data_ptr(), mask8.data_ptr(), N, C, V, H, W, dimg.data_ptr(), ops._st()))
        ops._count()
        return dimg, None, None


def pixel_gather(img, coord, mask):
    """img [V, C, H, W], coord [V, N, 2], mask bool [V, N] -> [N, C]: bilinear sample in the last camera that sees the point."""
    return _PixelGatherFn.apply(img, coord, mask)


def Feature_Gather(feature_map, xy, mode='bilinear'):
    """core/models/fusion_blocks.py:241-254 (every point sampled in every camera, no masks): [B, C, H, W], [B, N, 2] -> [B, C, N]."""
    assert mode == 'bilinear'
    B, C = feature_map.shape[:2]
    N = xy.shape[1]
    outs = []
    for b in range(B):
        m = torch.ones((1, N), dtype=torch.bool, device=xy.device)
        outs.append(pixel_gather(feature_map[b:b + 1], xy[b:b + 1], m).t())
    return torch.stack(outs, 0)


def Feature_Fetch(masks, pix_coord, imfeats, mode='bilinear'):
    """core/models/fusion_blocks.py:257-278: per batch element img [V, C, H, W], coord [V, N_b, 2], mask [V, N_b] -> [sum N_b, C]."""
    assert mode == 'bilinear'
    return torch.cat([pixel_gather(img, coord.to(img.device), mask.to(img.device)) for mask, coord, img in zip(masks, pix_coord, imfeats)], 0)
