"""Synthetic ScanNet-shaped scenes for tests and bench.py (there is no dataset in this repo).

Shape contract of the reference's input pipeline: (N, 4) fp32 = x, y, z, height with
height = z - 0.99-percentile-floor (mmdet3d/datasets/pipelines/loading.py:418-420), N = 40000
after IndoorPointSample (pipelines/transforms_3d.py:865-878); 18 ScanNet classes.  Geometry:
a room box x,y in [-4,4], z in [0,2.5]; 70 % of the points on floor / walls / 6-12 axis-aligned
cuboids (three quarters on their surfaces, one quarter inside, so some seeds lie near box centres) (so r = 0.2 balls hold tens of points, as in real scans), 30 % uniform;
0.5 % exact duplicates (exercises d2 == 0 in ball query and tie-breaking in FPS).
"""
import numpy as np
import torch

NUM_CLASSES = 18


def make_scene(seed, n_points=40000, dup_frac=0.005):
    """-> points (N,4) fp32, gt_boxes (G,7) fp32 [cx,cy,cz,sx,sy,sz,yaw=0], gt_labels (G,) int64."""
    rng = np.random.default_rng(1234 + seed)
    n_box = int(rng.integers(6, 13))
    size = rng.uniform(0.3, 1.8, (n_box, 3)).astype(np.float32)
    size[:, 2] = rng.uniform(0.3, 1.5, n_box)
    ctr = np.stack([rng.uniform(-3.2, 3.2, n_box), rng.uniform(-3.2, 3.2, n_box),
                    size[:, 2] / 2], axis=1).astype(np.float32)
    n_surf = int(n_points * 0.7)
    n_obj = int(n_surf * 0.55)
    # points on cuboid faces, proportional to face area
    which = rng.integers(0, n_box, n_obj)
    u = rng.uniform(-0.5, 0.5, (n_obj, 3)).astype(np.float32)
    face = rng.integers(0, 3, n_obj)
    sign = rng.choice([-0.5, 0.5], n_obj).astype(np.float32)
    interior = rng.uniform(0, 1, n_obj) < 0.25  # a quarter of the object points fill the volume
    u[np.arange(n_obj), face] = np.where(interior, u[np.arange(n_obj), face], sign)
    obj_pts = ctr[which] + u * size[which]
    # floor and walls
    n_floor = (n_surf - n_obj) // 2
    n_wall = n_surf - n_obj - n_floor
    floor = np.stack([rng.uniform(-4, 4, n_floor), rng.uniform(-4, 4, n_floor),
                      rng.normal(0, 0.01, n_floor)], axis=1)
    wall = np.stack([rng.uniform(-4, 4, n_wall), rng.uniform(-4, 4, n_wall),
                     rng.uniform(0, 2.5, n_wall)], axis=1)
    side = rng.integers(0, 4, n_wall)
    wall[side == 0, 0], wall[side == 1, 0] = -4.0, 4.0
    wall[side == 2, 1], wall[side == 3, 1] = -4.0, 4.0
    n_uni = n_points - n_surf
    uni = np.stack([rng.uniform(-4, 4, n_uni), rng.uniform(-4, 4, n_uni),
                    rng.uniform(0, 2.5, n_uni)], axis=1)
    xyz = np.concatenate([obj_pts, floor, wall, uni], axis=0).astype(np.float32)
    xyz = xyz[rng.permutation(n_points)]
    n_dup = int(n_points * dup_frac)
    if n_dup:
        dst = rng.choice(n_points, n_dup, replace=False)
        src = rng.choice(n_points, n_dup, replace=False)
        xyz[dst] = xyz[src]
    floor_h = np.percentile(xyz[:, 2], 0.99)
    pts = np.concatenate([xyz, (xyz[:, 2:3] - floor_h)], axis=1).astype(np.float32)
    boxes = np.concatenate([ctr, size, np.zeros((n_box, 1), np.float32)], axis=1).astype(np.float32)
    labels = rng.integers(0, NUM_CLASSES, n_box).astype(np.int64)
    return torch.from_numpy(pts), torch.from_numpy(boxes), torch.from_numpy(labels)


def make_batch(batch, n_points=40000, seed0=0, origin='gravity'):
    """-> points (B,N,4) fp32 (CPU), list of gt boxes, list of gt labels.  origin='bottom': boxes as
    DepthInstance3DBoxes stores them (z of the bottom face; what NesieHead.loss expects)."""
    scenes = [make_scene(seed0 + i, n_points) for i in range(batch)]
    if origin == 'bottom':
        for s in scenes:
            s[1][:, 2] -= 0.5 * s[1][:, 5]
    return (torch.stack([s[0] for s in scenes]), [s[1] for s in scenes], [s[2] for s in scenes])
