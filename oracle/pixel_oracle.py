"""CPU oracle of the student's point <-> pixel transforms — TEST INFRASTRUCTURE.  The reference code IS plain torch, so the
oracle is its statement-by-statement restatement (device-agnostic: `.cuda()` calls dropped).  PINNED to the reference's own
code: tests/test_pixel_oracle_pin_cpu.py runs core/models/fusion_blocks.py's Feature_Gather / Feature_Fetch and the unmodified
student model's inline multi-scale loop (captured with forward hooks) on CPU next to these functions — difference exactly 0.
  Point2Grid      core/models/fusion_blocks.py:217-238
  Feature_Gather  core/models/fusion_blocks.py:241-254
  Feature_Fetch   core/models/fusion_blocks.py:257-278
  multiscale_point2grid   core/models/nuscenes/spvcnn_swiftnet18_spformer_tsd_full.py:448-478"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def Point2Grid(pts_feat, pixel_coordinates, masks, grid_size):
    cur, l2c_feat_map = 0, []
    h, w = grid_size
    for mask, coord in zip(masks, pixel_coordinates):
        n = mask.size(1)
        bs_pts_feat = pts_feat[cur:cur + n, :]
        for co, ma in zip(coord, mask):
            u = (co[:, 0] + 1.0) / 2 * (w - 1.0)
            v = (co[:, 1] + 1.0) / 2 * (h - 1.0)
            uv = torch.floor(torch.stack([u, v], dim=1)).long()
            uv = torch.fliplr(uv[ma])
            uq, inv, count = torch.unique(uv, dim=0, return_inverse=True, return_counts=True)
            f2d = torch.zeros(size=(uq.size(0), pts_feat.size(1)), dtype=pts_feat.dtype)
            inv = inv.view(-1, 1).expand(-1, f2d.size(-1))
            f2d = f2d.scatter_add(0, inv, bs_pts_feat[ma])
            f2d = f2d / count.view(-1, 1)
            l2c_f = torch.sparse_coo_tensor(uq.transpose(0, 1).contiguous(), f2d, size=(h, w, f2d.size(-1)))
            l2c_feat_map.append(l2c_f.to_dense())
        cur += n
    return torch.stack(l2c_feat_map, dim=0).permute(0, 3, 1, 2).contiguous()


def multiscale_point2grid(pts_feat, pixel_coordinates, masks, grid_size, n_scales):
    ifh, ifw = grid_size
    cur, l2c_feat_map = 0, []
    for mask, coord in zip(masks, pixel_coordinates):
        n = mask.size(1)
        bs_pts_feat = pts_feat[cur:cur + n, :]
        for co, ma in zip(coord, mask):
            l2c_f = torch.zeros(size=(1, pts_feat.size(1), ifh, ifw), dtype=pts_feat.dtype)
            if torch.sum(ma) == 0:
                l2c_feat_map.append(l2c_f / n_scales)
                continue
            cnt = 1
            for _ in range(n_scales):
                c_ih = int(round(float(ifh) / cnt + 0.01))
                c_iw = int(round(float(ifw) / cnt + 0.01))
                u = (co[:, 0] + 1.0) / 2 * (c_iw - 1.0)
                v = (co[:, 1] + 1.0) / 2 * (c_ih - 1.0)
                uv = torch.floor(torch.stack([u, v], dim=1)).long()
                uv = torch.fliplr(uv[ma])
                uq, inv, count = torch.unique(uv, dim=0, return_inverse=True, return_counts=True)
                f2d = torch.zeros(size=(uq.size(0), pts_feat.size(1)), dtype=pts_feat.dtype)
                inv = inv.view(-1, 1).expand(-1, f2d.size(-1))
                f2d = f2d.scatter_add(0, inv, bs_pts_feat[ma])
                f2d = f2d / count.view(-1, 1)
                tmp = torch.sparse_coo_tensor(uq.transpose(0, 1).contiguous(), f2d, size=(c_ih, c_iw, f2d.size(-1))
                                              ).to_dense().permute(2, 0, 1).contiguous().view(1, -1, c_ih, c_iw)
                l2c_f = l2c_f + F.interpolate(tmp, size=(ifh, ifw), mode='bilinear', align_corners=True)  # build_blocks.py:18
                cnt *= 2
            l2c_feat_map.append(l2c_f / n_scales)
        cur += n
    return torch.concat(l2c_feat_map, dim=0).contiguous()


def Feature_Gather(feature_map, xy, mode='bilinear'):
    xy = xy.unsqueeze(1)
    return nn.functional.grid_sample(feature_map, xy, padding_mode='zeros', align_corners=True, mode=mode).squeeze(2)


def Feature_Fetch(masks, pix_coord, imfeats, mode='bilinear'):
    imfs = []
    for mask, coord, img in zip(masks, pix_coord, imfeats):
        imf = torch.zeros(size=(mask.size(1), img.size(1)), dtype=img.dtype)
        imf_list = Feature_Gather(img, coord, mode=mode).permute(0, 2, 1)
        for idx in range(mask.size(0)):
            imf = imf.clone()
            imf[mask[idx]] = imf_list[idx, mask[idx], :]
        imfs.append(imf)
    return torch.cat(imfs, dim=0)
