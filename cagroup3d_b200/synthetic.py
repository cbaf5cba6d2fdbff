"""Synthetic ScanNet- / SUN RGB-D-shaped scenes and the reference `collate_batch` layout.

Host-side numpy, used by bench.py, tools/test.py (synthetic dataset) and the tests.  A scene is a room
(floor + 4 walls) with 12 boxes resting on the floor, sampled on surfaces with 4 mm noise; points are
drawn until the 0.02 m voxel count reaches the target.
"""
from __future__ import annotations

import numpy as np

from . import model_init


def _box_surface(rng, n, center, size):
    """n points uniform on the surface of an axis-aligned box."""
    areas = np.array([size[1] * size[2], size[0] * size[2], size[0] * size[1]] * 2)
    face = rng.choice(6, size=n, p=areas / areas.sum())
    p = (rng.random((n, 3)) - 0.5) * size
    ax = face % 3
    sign = np.where(face < 3, -0.5, 0.5)
    p[np.arange(n), ax] = sign * size[ax]
    return p + center


def make_scene(seed: int, target_voxels: int = 50000, voxel_size: float = 0.02, n_classes: int = 18,
               sunrgbd: bool = False, n_points: int | None = None, boxes_only: bool = False, return_masks: bool = False):
    """-> (points (N,6) f32 [x,y,z,r,g,b] colours 0..255, gt_boxes (12,8) [x,y,z,dx,dy,dz,yaw,cls]); boxes_only: just the
    boxes (they are drawn first from the scene's seed, so they equal the boxes of the full scene).  return_masks: also the
    per-point (semantic_mask, instance_mask) int64 arrays of the ScanNet training items (scannet_dataset.py:68-76): a
    point sampled on box i has instance 5 + i and the box's class, floor / wall points have instances 0..4 and the
    background class n_classes.  The masks are derived from the part counts, the random stream is unchanged."""
    rng = np.random.default_rng(seed)
    L, W, H = rng.uniform(4, 8), rng.uniform(3, 6), 2.6
    boxes = []
    for i in range(12):
        sz = np.array([rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.2), rng.uniform(0.4, 1.2)])
        c = np.array([rng.uniform(-L / 2 + sz[0] / 2, L / 2 - sz[0] / 2), rng.uniform(-W / 2 + sz[1] / 2, W / 2 - sz[1] / 2),
                      -H / 2 + sz[2] / 2])
        yaw = rng.uniform(-np.pi, np.pi) if sunrgbd else 0.0
        boxes.append(np.concatenate([c, sz, [yaw, i % n_classes]]))
    boxes = np.array(boxes, dtype=np.float32)
    if boxes_only:
        return boxes

    def sample(n):
        parts = []
        counts = rng.multinomial(n, [0.30] + [0.10] * 4 + [0.025] * 12)
        f = (rng.random((counts[0], 3)) - 0.5) * [L, W, 0]
        f[:, 2] = -H / 2
        parts.append(f)
        for w, cnt in enumerate(counts[1:5]):
            q = (rng.random((cnt, 3)) - 0.5) * [L, W, H]
            if w < 2:
                q[:, 0] = (-L / 2, L / 2)[w]
            else:
                q[:, 1] = (-W / 2, W / 2)[w - 2]
            parts.append(q)
        for b, cnt in zip(boxes, counts[5:]):
            q = _box_surface(rng, cnt, np.zeros(3), b[3:6].astype(np.float64))
            if sunrgbd:
                c_, s_ = np.cos(b[6]), np.sin(b[6])
                q[:, :2] = q[:, :2] @ np.array([[c_, s_], [-s_, c_]])
            parts.append(q + b[:3])
        p = np.concatenate(parts)
        part_of.append(np.repeat(np.arange(17), counts))
        return p + rng.normal(0, 0.004, p.shape)

    part_of = []
    pts = sample(int(target_voxels * 1.05))
    for _ in range(40):
        nv = len(np.unique(np.floor(pts / voxel_size).astype(np.int64), axis=0))
        if nv >= target_voxels:
            break
        pts = np.concatenate([pts, sample(max(256, int((target_voxels - nv) * 1.3)))])
    part = np.concatenate(part_of)
    if sunrgbd:
        ang = np.arctan2(pts[:, 1], pts[:, 0] + L / 2 + 0.5)
        sel = np.abs(ang) < np.pi / 6
        pts, part = pts[sel], part[sel]
    if n_points is not None:
        idx = rng.choice(len(pts), n_points, replace=len(pts) < n_points)
        pts, part = pts[idx], part[idx]
    rgb = rng.integers(0, 256, (len(pts), 3)).astype(np.float32)
    out = np.concatenate([pts.astype(np.float32), rgb], 1)
    if return_masks:
        sem = np.where(part >= 5, boxes[np.clip(part - 5, 0, 11), 7].astype(np.int64), n_classes)
        return out, boxes, sem.astype(np.int64), part.astype(np.int64)
    return out, boxes


def collate_batch(scenes):
    """pcdet/datasets/dataset.py:172-177: prepend the sample index -> points (sum N, 7)."""
    pts = [np.concatenate([np.full((len(p), 1), i, np.float32), p], 1) for i, (p, _) in enumerate(scenes)]
    gt = np.stack([b for _, b in scenes])
    return {"points": np.concatenate(pts), "gt_boxes": gt, "batch_size": len(scenes),
            "frame_id": np.arange(len(scenes))}


def make_batch(batch_size: int, target_voxels: int = 50000, config: int = 2, n_classes: int = 18,
               sunrgbd: bool = False, first_scene: int = 0, n_points: int | None = None):
    """seed = 1000 * config + scene index."""
    return collate_batch([make_scene(1000 * config + first_scene + i, target_voxels, n_classes=n_classes,
                                     sunrgbd=sunrgbd, n_points=n_points) for i in range(batch_size)])


def model_cfg(n_classes: int = 18, with_yaw: bool = False) -> dict:
    """The MODEL section of tools/cfgs/{scannet,sunrgbd}_models/CAGroup3D.yaml as a plain dict."""
    return model_init.default_model_cfg(n_classes, with_yaw)
