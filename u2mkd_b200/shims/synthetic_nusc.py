"""Synthetic stand-in for core/datasets/semantic_nusc.py's NuScenes: same feed_dict, no dataset on disk.

Each item follows semantic_nusc.py:258-351 statement by statement — keyframe + aggregated sweeps with a keyframe mask and
ignored labels on the sweep points, flip / rotate-scale / translate augmentation on the train split, round(xyz / voxel_size)
voxel coordinates shifted to start at 0, sparse_quantize (first point per voxel) — with the points coming from
u2mkd_b200.scans (ray-cast boxes on a ground plane, nuScenes sensor layout) and labels from geometry (ground / box / far)
instead of .bin / lidarseg files.  Keys and types:

    lidar               SparseTensor(feats fp32 [n, 4] (x, y, z, intensity), coords int32 [n, 3])   first point per voxel
    targets             SparseTensor(labels [n], coords)
    targets_mapped      SparseTensor(labels of ALL points [N], voxel coords of all points [N, 3])
    inverse_map         SparseTensor(inverse [N], voxel coords of all points)
    lidar_token         str
    num_vox             int
    keyframe_mask       SparseTensor(bool [n], coords)            (multisweeps != 0)
    keyframe_mask_full  SparseTensor(bool [N], all voxel coords)  (multisweeps != 0)

`collate_fn` is the reference's (semantic_nusc.py:353-376): SparseTensors through sparse_collate (batch index appended),
everything else listed.  The SparseTensor / collate come from the `torchsparse` that is installed at call time
(u2mkd_b200.install_as_torchsparse(), or the oracle's namespace in the CPU tests)."""
from __future__ import annotations

import numpy as np

from .. import scans


class SyntheticNuScenesSplit:
    def __init__(self, split: str, voxel_size: float = 0.1, num_samples: int = 8, multisweeps: int = 0, num_classes: int = 17,
                 ignored_label: int = 0, seed: int = 0, flip_aug: bool = True, rotate_aug: bool = True,
                 translate_std=(0.1, 0.1, 0.1), max_points: int = 0):
        self.split, self.voxel_size, self.num_samples = split, voxel_size, num_samples
        self.multisweeps, self.num_classes, self.ignored_labels = multisweeps, num_classes, ignored_label
        self.seed, self.flip_aug, self.rotate_aug, self.translate_std = seed, flip_aug, rotate_aug, translate_std
        self.max_points = max_points

    def __len__(self) -> int:
        return self.num_samples

    def _labels(self, pts: np.ndarray) -> np.ndarray:
        """Geometry classes in [1, num_classes): ground by height, the rest by range ring and azimuth sector."""
        r = np.hypot(pts[:, 0], pts[:, 1])
        ground = pts[:, 2] < (pts[:, 2].min() + 0.3)
        sector = ((np.arctan2(pts[:, 1], pts[:, 0]) + np.pi) / (2 * np.pi) * 4).astype(np.int64) % 4
        ring = np.minimum((r / 15.0).astype(np.int64), 2)
        lab = 1 + (sector * 3 + ring) % (self.num_classes - 2)
        lab[ground] = self.num_classes - 1
        return lab.astype(np.uint8)

    def __getitem__(self, index: int):
        from torchsparse import SparseTensor
        from torchsparse.utils.quantize import sparse_quantize
        train = "train" in self.split
        rng_state = np.random.get_state() if not train else None
        seed = self.seed * 100003 + (0 if train else 50000) + index
        pts = scans.raw_scan(seed, "nusc", 1)                                   # keyframe [N, 4]
        if self.max_points and pts.shape[0] > self.max_points:
            pts = pts[np.random.default_rng(seed).permutation(pts.shape[0])[:self.max_points]]
        labels_ = self._labels(pts)
        if self.multisweeps != 0:                                                # semantic_nusc.py:281-289
            full = scans.raw_scan(seed, "nusc", 1 + self.multisweeps)
            extra = full[scans.raw_scan(seed, "nusc", 1).shape[0]:]
            if self.max_points:
                extra = extra[:self.max_points // 2]
            agg_ts = np.concatenate([np.zeros(pts.shape[0]), np.full(extra.shape[0], 0.05)])
            keyframe_mask = agg_ts == 0
            labels_ = np.concatenate([labels_, np.full(int((~keyframe_mask).sum()), self.ignored_labels, np.uint8)])
            pts = np.concatenate([pts, extra], 0)
        pts = pts.astype(np.float32).copy()
        if train and self.flip_aug:                                              # :291-298
            flip_type = np.random.choice(4, 1)
            if flip_type == 1:
                pts[:, 0] = -pts[:, 0]
            elif flip_type == 2:
                pts[:, 1] = -pts[:, 1]
            elif flip_type == 3:
                pts[:, :2] = -pts[:, :2]
        pts_cp = np.zeros_like(pts)
        if train and self.rotate_aug:                                            # :300-315
            theta = np.random.uniform(0, 2 * np.pi)
            scale_factor = np.random.uniform(0.95, 1.05)
            rot = np.array([[np.cos(theta), np.sin(theta), 0], [-np.sin(theta), np.cos(theta), 0], [0, 0, 1]])
            pts_cp[:, :3] = np.dot(pts[:, :3], rot) * scale_factor
        else:
            pts_cp[...] = pts[...]
        if train and self.translate_std:                                         # :317-322
            pts_cp[:, :3] += np.array([np.random.normal(0, s, 1) for s in self.translate_std]).T
        pts_cp[:, 3] = pts[:, 3]
        voxel = np.round(pts_cp[:, :3] / self.voxel_size).astype(np.int32)       # :324-326
        voxel -= voxel.min(0, keepdims=1)
        feat_ = pts_cp.astype(np.float32)
        _, inds, inverse_map = sparse_quantize(voxel, return_index=True, return_inverse=True)
        voxel_full, feat_full, labels_full = voxel[inds], feat_[inds], labels_[inds]
        feed_dict = {
            "lidar": SparseTensor(feat_full, voxel_full),
            "targets": SparseTensor(labels_full, voxel_full),
            "targets_mapped": SparseTensor(labels_, voxel),
            "inverse_map": SparseTensor(inverse_map, voxel),
            "lidar_token": f"synthetic-{self.split}-{index:06d}",
            "num_vox": voxel_full.shape[0],
        }
        if self.multisweeps != 0:
            feed_dict["keyframe_mask"] = SparseTensor(keyframe_mask[inds], voxel_full)
            feed_dict["keyframe_mask_full"] = SparseTensor(keyframe_mask, voxel)
        if rng_state is not None:
            np.random.set_state(rng_state)
        return feed_dict

    @staticmethod
    def collate_fn(batch):
        import torch
        from torchsparse import SparseTensor
        from torchsparse.utils.collate import sparse_collate, sparse_collate_fn
        if not isinstance(batch[0], dict):
            return batch
        out = {}
        for key, first in batch[0].items():
            col = [sample[key] for sample in batch]
            if isinstance(first, SparseTensor):
                out[key] = sparse_collate(col)
            elif isinstance(first, np.ndarray):
                out[key] = torch.stack([torch.from_numpy(v).float() for v in col], dim=0)
            elif isinstance(first, torch.Tensor):
                out[key] = torch.stack(col, dim=0)
            elif isinstance(first, dict):
                out[key] = sparse_collate_fn(col)
            else:
                out[key] = col
        return out


class SyntheticNuScenes(dict):
    """{'train': split, 'val': split} — what builder.make_dataset() returns for `semantic_nusc` (semantic_nusc.py:27-60)."""

    def __init__(self, voxel_size: float = 0.1, num_train: int = 8, num_val: int = 4, **kw):
        super().__init__({
            "train": SyntheticNuScenesSplit("train", voxel_size, num_train, **kw),
            "val": SyntheticNuScenesSplit("val", voxel_size, num_val, **kw),
        })


# ------------------------------------------------------------------------------------------ LiDAR + six cameras (student)
class SyntheticNuScenesCamerasSplit(SyntheticNuScenesSplit):
    """core/datasets/lc_semantic_nusc_tsd_full.py:313-462 without the files: every item is
        {'feed_dict_s': student input, 'feed_dict_t': teacher input, 'lidar_token': str}
    feed_dict_t (`_process_unimodal_input`, :191-240): the keyframe plus the aggregated sweeps, own rotate / scale / flip
        augmentation — lidar, targets, targets_mapped, inverse_map, num_vox, num_pts (+ keyframe_mask, keyframe_mask_full);
        the keyframe points come FIRST and in the student's order (core/nusc_trainers.py:292-302 maps teacher logits to the
        student's voxels through inverse_map -> keyframe_mask_full -> the student's `inds`).
    feed_dict_s: the keyframe only — lidar, targets, targets_mapped, inverse_map, num_vox, inds ([first-point indices]),
        images fp32 [V, H, W, 3] (V = 6 cameras, minus `im_drop` random ones on the train split), pixel_coordinates fp32
        [V, n, 2] ((u, v) in [-1, 1], u along the image width), masks bool [V, n] (in front of the camera and inside the
        image), fov_mask SparseTensor(seen by any camera), label_fov (debug_val: labels of the points some camera sees,
        ignore elsewhere).
    Cameras: six pinholes at the nuScenes yaw layout, 1.5 m above the ground; images are seeded noise (the camera branch
    only needs the shapes)."""

    CAM_YAW_DEG = (55.0, 0.0, -55.0, 110.0, 180.0, -110.0)   # FRONT_LEFT, FRONT, FRONT_RIGHT, BACK_LEFT, BACK, BACK_RIGHT

    def __init__(self, split, voxel_size=0.1, num_samples=8, image_size=(360, 640), im_drop=3, debug=True, **kw):
        super().__init__(split, voxel_size, num_samples, **kw)
        self.image_size, self.im_drop, self.debug = tuple(image_size), im_drop, debug

    def _rotate_and_scale(self, pts):
        theta = np.random.uniform(0, 2 * np.pi)
        scale_factor = np.random.uniform(0.95, 1.05)
        rot = np.array([[np.cos(theta), np.sin(theta), 0], [-np.sin(theta), np.cos(theta), 0], [0, 0, 1]])
        out = np.zeros_like(pts)
        out[:, :3] = np.dot(pts[:, :3], rot) * scale_factor
        out[:, 3] = pts[:, 3]
        return out

    def _quantize(self, pts_cp, labels_):
        from torchsparse import SparseTensor
        from torchsparse.utils.quantize import sparse_quantize
        voxel = np.round(pts_cp[:, :3] / self.voxel_size).astype(np.int32)
        voxel -= voxel.min(0, keepdims=1)
        feat_ = pts_cp.astype(np.float32)
        _, inds, inverse_map = sparse_quantize(voxel, return_index=True, return_inverse=True)
        voxel_full = voxel[inds]
        d = {"lidar": SparseTensor(feat_[inds], voxel_full), "targets": SparseTensor(labels_[inds], voxel_full),
             "targets_mapped": SparseTensor(labels_, voxel), "inverse_map": SparseTensor(inverse_map, voxel),
             "num_vox": voxel_full.shape[0]}
        return d, voxel, voxel_full, inds

    def __getitem__(self, index: int):
        from torchsparse import SparseTensor
        train = "train" in self.split
        rng_state = np.random.get_state() if not train else None
        seed = self.seed * 100003 + (0 if train else 50000) + index
        rng = np.random.default_rng(seed)
        key = scans.raw_scan(seed, "nusc", 1)
        if self.max_points and key.shape[0] > self.max_points:
            key = key[rng.permutation(key.shape[0])[:self.max_points]]
        pts = key.astype(np.float32)
        labels_raw = self._labels(pts)

        # ---- teacher: keyframe + sweeps (:191-240)
        t_pts, t_labels = pts.copy(), labels_raw
        keyframe_mask = None
        if self.multisweeps != 0:
            extra = scans.raw_scan(seed, "nusc", 1 + self.multisweeps)[scans.raw_scan(seed, "nusc", 1).shape[0]:]
            if self.max_points:
                extra = extra[:self.max_points // 2]
            keyframe_mask = np.concatenate([np.ones(pts.shape[0], bool), np.zeros(extra.shape[0], bool)])
            t_labels = np.concatenate([labels_raw, np.full(extra.shape[0], self.ignored_labels, np.uint8)])
            t_pts = np.concatenate([t_pts, extra.astype(np.float32)], 0)
        if train:
            t_pts = self._rotate_and_scale(t_pts)
            flip_type = np.random.choice(4, 1)
            if flip_type == 1:
                t_pts[:, 0] = -t_pts[:, 0]
            elif flip_type == 2:
                t_pts[:, 1] = -t_pts[:, 1]
            elif flip_type == 3:
                t_pts[:, :2] = -t_pts[:, :2]
        feed_t, t_voxel, t_voxel_full, t_inds = self._quantize(t_pts, t_labels)
        feed_t["num_pts"] = t_voxel.shape[0]
        if keyframe_mask is not None:
            feed_t["keyframe_mask"] = SparseTensor(keyframe_mask[t_inds], t_voxel_full)
            feed_t["keyframe_mask_full"] = SparseTensor(keyframe_mask, t_voxel)

        # ---- student: cameras over the un-augmented keyframe (:330-392)
        h, w = self.image_size
        focal = 0.55 * w
        cams = list(range(6))
        if train and self.im_drop:
            drop = set(np.random.choice(6, self.im_drop, replace=False).tolist())
            cams = [c for c in cams if c not in drop]
        images, pixel_coordinates, masks = [], [], []
        valid_mask = np.full(pts.shape[0], -1)
        ground = float(pts[:, 2].min())
        for idx in cams:
            yaw = np.deg2rad(self.CAM_YAW_DEG[idx])
            fwd = pts[:, 0] * np.cos(yaw) + pts[:, 1] * np.sin(yaw)          # depth along the optical axis
            left = -pts[:, 0] * np.sin(yaw) + pts[:, 1] * np.cos(yaw)
            up = pts[:, 2] - (ground + 1.5)
            mask = fwd > 1
            depth = np.where(mask, fwd, 1.0)
            u = (-left / depth * focal + (w - 1) / 2) / (w - 1.0) * 2.0 - 1.0   # width
            v = (-up / depth * focal + (h - 1) / 2) / (h - 1.0) * 2.0 - 1.0     # height
            mask = mask & (u > -1) & (u < 1) & (v > -1) & (v < 1)
            valid_mask[mask] = idx
            masks.append(mask)
            pixel_coordinates.append(np.stack([u, v], 1))
            images.append(rng.standard_normal((h, w, 3)).astype(np.float32))
        pt_with_img_idx = valid_mask != -1
        pixel_coordinates = np.stack(pixel_coordinates, 0)
        masks = np.stack(masks, 0)
        images = np.stack(images, 0)
        pts_cp = self._rotate_and_scale(pts) if train else pts.copy()            # (:398-414)
        feed_s, voxel, voxel_full, inds = self._quantize(pts_cp, labels_raw)
        feed_s.update({
            "images": images,
            "pixel_coordinates": pixel_coordinates[:, inds, :].astype(np.float32),
            "masks": masks[:, inds],
            "fov_mask": SparseTensor(pt_with_img_idx[inds], voxel_full),
            "inds": [inds],
        })
        if self.debug:
            label_fov = np.full_like(labels_raw, fill_value=self.ignored_labels, dtype=np.uint8)
            label_fov[pt_with_img_idx] = labels_raw[pt_with_img_idx]
            feed_s["label_fov"] = SparseTensor(label_fov, voxel)
        if rng_state is not None:
            np.random.set_state(rng_state)
        return {"feed_dict_s": feed_s, "feed_dict_t": feed_t, "lidar_token": f"synthetic-{self.split}-{index:06d}"}

    @staticmethod
    def collate_fn(batch):
        """lc_semantic_nusc_tsd_full.py:464-488: masks / pixel_coordinates stay per-sample lists, nested dicts recurse."""
        import torch
        from torchsparse import SparseTensor
        from torchsparse.utils.collate import sparse_collate
        if not isinstance(batch[0], dict):
            return batch
        out = {}
        for key, first in batch[0].items():
            col = [sample[key] for sample in batch]
            if key == "masks":
                out[key] = [torch.from_numpy(v) for v in col]
            elif key == "pixel_coordinates":
                out[key] = [torch.from_numpy(v).float() for v in col]
            elif isinstance(first, SparseTensor):
                out[key] = sparse_collate(col)
            elif isinstance(first, np.ndarray):
                out[key] = torch.stack([torch.from_numpy(v).float() for v in col], dim=0)
            elif isinstance(first, torch.Tensor):
                out[key] = torch.stack(col, dim=0)
            elif isinstance(first, dict):
                out[key] = SyntheticNuScenesCamerasSplit.collate_fn(col)
            else:
                out[key] = col
        return out


class SyntheticNuScenesCameras(dict):
    """What builder.make_dataset() returns for `lc_semantic_nusc_tsd_full` (lc_semantic_nusc_tsd_full.py:64-71)."""

    def __init__(self, voxel_size: float = 0.1, num_train: int = 8, num_val: int = 4, **kw):
        super().__init__({
            "train": SyntheticNuScenesCamerasSplit("train", voxel_size, num_train, **kw),
            "val": SyntheticNuScenesCamerasSplit("val", voxel_size, num_val, **kw),
        })
