"""Seeded synthetic LiDAR scans shaped like the reference's datasets.

Emits exactly the input contract of the hot path
(/root/reference/core/datasets/semantic_nusc.py:319-336 + sparse_collate):
    voxel = round(xyz / voxel_size).int32 ; voxel -= voxel.min(0)
    first point per voxel (sorted by ravel hash)  -> coords int32 [N,3], feats fp32 [N,4]
and, after collation, coords int32 [N,4] = (x, y, z, batch) with the batch index LAST.
Pure numpy; no datasets are read.
"""
from __future__ import annotations

import numpy as np

NUSC = dict(beams=32, elev=(-30.67, 10.67), azimuth=1085, height=1.84, max_range=70.0)
KITTI = dict(beams=64, elev=(-24.8, 2.0), azimuth=2083, height=1.73, max_range=80.0)


def _scene(rng, n_boxes=25, extent=60.0):
    """Axis-aligned boxes (cx, cy, cz, hx, hy, hz): vehicles + two long walls."""
    boxes = []
    for _ in range(n_boxes):
        cx, cy = rng.uniform(-extent, extent, 2)
        if abs(cx) < 3 and abs(cy) < 3:
            cx += 6.0
        hx, hy, hz = rng.uniform(0.8, 2.6), rng.uniform(0.8, 2.6), rng.uniform(0.7, 1.6)
        boxes.append((cx, cy, hz, hx, hy, hz))
    boxes.append((0.0, 18.0 + rng.uniform(-3, 3), 3.0, 80.0, 0.3, 3.0))
    boxes.append((0.0, -22.0 + rng.uniform(-3, 3), 4.0, 80.0, 0.3, 4.0))
    return np.asarray(boxes, dtype=np.float64)


def _raycast(origin, dirs, boxes, max_range):
    """Nearest hit of each ray with the ground plane z=0 and the boxes (slab test)."""
    t_best = np.full(dirs.shape[0], np.inf)
    dz = dirs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(dz < -1e-9, -origin[2] / dz, np.inf)
    t_best = np.minimum(t_best, tg)
    inv = 1.0 / np.where(np.abs(dirs) < 1e-12, 1e-12, dirs)
    for b in boxes:
        lo = (b[:3] - b[3:] - origin) * inv
        hi = (b[:3] + b[3:] - origin) * inv
        tmin = np.minimum(lo, hi).max(axis=1)
        tmax = np.maximum(lo, hi).min(axis=1)
        hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.3)
        t_best = np.where(hit & (tmin < t_best), tmin, t_best)
    ok = np.isfinite(t_best) & (t_best < max_range)
    return t_best, ok


def raw_sweep(rng, sensor, boxes, ego_shift=0.0):
    """One spinning-LiDAR sweep -> float64 [n,4] = (x, y, z, intensity) in the keyframe's frame."""
    el = np.deg2rad(np.linspace(sensor["elev"][0], sensor["elev"][1], sensor["beams"]))
    az = np.linspace(0, 2 * np.pi, sensor["azimuth"], endpoint=False) + rng.uniform(0, 2 * np.pi / sensor["azimuth"])
    el_g, az_g = np.meshgrid(el, az, indexing="ij")
    dirs = np.stack([np.cos(el_g) * np.cos(az_g), np.cos(el_g) * np.sin(az_g), np.sin(el_g)], -1).reshape(-1, 3)
    origin = np.array([ego_shift, 0.0, sensor["height"]])
    t, ok = _raycast(origin, dirs, boxes, sensor["max_range"])
    keep = ok & (rng.uniform(size=ok.shape) > 0.04)
    pts = origin + dirs[keep] * t[keep, None]
    pts += rng.normal(0, 0.02, pts.shape)
    inten = rng.uniform(0, 255, (pts.shape[0], 1))
    return np.concatenate([pts, inten], 1)


def raw_scan(seed: int, kind: str = "nusc", sweeps: int = 1):
    """Keyframe plus (sweeps-1) extra sweeps with 0.5 m ego shift each; near-ego points of the
    extra sweeps are dropped like semantic_nusc.py:172-175,189."""
    rng = np.random.default_rng(seed)
    sensor = NUSC if kind == "nusc" else KITTI
    boxes = _scene(rng)
    out = [raw_sweep(rng, sensor, boxes, 0.0)]
    for s in range(1, sweeps):
        shift = 0.5 * ((s + 1) // 2) * (1 if s % 2 else -1)
        p = raw_sweep(rng, sensor, boxes, shift)
        near = (np.abs(p[:, 0] - shift) < 1.0) & (np.abs(p[:, 1]) < 1.0)
        out.append(p[~near])
    return np.concatenate(out, 0).astype(np.float32)


def quantize_scan(pts: np.ndarray, voxel_size: float):
    """Dataset contract: returns (coords int32 [n,3], feats fp32 [n,4], inds, inverse)."""
    voxel = np.round(pts[:, :3] / voxel_size).astype(np.int32)
    voxel -= voxel.min(0, keepdims=True)
    v = voxel.astype(np.uint64)
    vmax = v.max(0) + 1
    key = (v[:, 0] * vmax[1] + v[:, 1]) * vmax[2] + v[:, 2]
    _, inds, inverse = np.unique(key, return_index=True, return_inverse=True)
    return voxel[inds], pts[inds].astype(np.float32), inds, inverse


def make_batch(seeds, kind="nusc", sweeps=1, voxel_size=0.1):
    """Collated batch: coords int32 [N,4] (x,y,z,batch), feats fp32 [N,4]; one scan per seed."""
    cs, fs = [], []
    for b, seed in enumerate(seeds):
        c, f, _, _ = quantize_scan(raw_scan(seed, kind, sweeps), voxel_size)
        cs.append(np.concatenate([c, np.full((c.shape[0], 1), b, np.int32)], 1))
        fs.append(f)
    return np.ascontiguousarray(np.concatenate(cs, 0)), np.ascontiguousarray(np.concatenate(fs, 0))


WORKLOADS = {
    # BASELINE.json configs[0]: 1-sweep nuScenes-shape scan, 0.1 m voxels, cr 0.5
    "nusc1_cr0.5": dict(kind="nusc", sweeps=1, voxel_size=0.1, cr=0.5, batch=1),
    # BASELINE.json configs[1]/[2]: multisweep nuScenes-shape scan (~140 k pts), 0.05 m, cr 2.0, batch 2
    "nusc5_cr2.0_b2": dict(kind="nusc", sweeps=5, voxel_size=0.05, cr=2.0, batch=2),
    # BASELINE.json configs[3]: SemanticKITTI-shape scan, 0.05 m
    "kitti1": dict(kind="kitti", sweeps=1, voxel_size=0.05, cr=1.0, batch=1),
}
