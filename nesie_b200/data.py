"""On-disk format and input pipeline of the path.

Mirror of the reference's ScanNet input pipeline, the part that feeds the hot path:
  * LoadPointsFromFile (datasets/pipelines/loading.py:366-420): `.bin` = raw little-endian float32,
    `load_dim` values per point, columns `use_dim`; `shift_height` appends height = z - the 0.99th
    percentile of z (numpy's linear-interpolation percentile) as the 4th feature;
  * IndoorPointSample (pipelines/transforms_3d.py:821-891): every scene to exactly num_points points,
    without replacement when the scene has enough points, with replacement otherwise;
  * RandomFlip3D / GlobalRotScaleTrans: the augmentation record `transformation_3d_flow` that
    VoteNetNesie.transformation_bbox_preds replays on the teacher's boxes (detectors.BoxAug).
The reference does all of this in 4 CPU worker processes per GPU with numpy; at B200 step rates that
is the bottleneck, so here the file bytes go to pinned memory and everything after the percentile
runs on the device: sampling is one randperm / randint + gather per scene, the two augmented views
of a mean-teacher batch are produced by BoxAug.apply_points for the whole batch at once.
"""
import numpy as np
import torch

from .detectors import BoxAug, transform_boxes


def points_from_bytes(buf, load_dim=6, use_dim=(0, 1, 2), shift_height=True):
    """Raw `.bin` bytes (or a float32 array) -> (N, len(use_dim) [+ 1]) float32 numpy array, exactly
    LoadPointsFromFile.__call__ (loading.py:411-420)."""
    points = np.frombuffer(buf, dtype=np.float32) if isinstance(buf, (bytes, bytearray, memoryview)) \
        else np.asarray(buf, dtype=np.float32).reshape(-1)
    points = points.reshape(-1, load_dim)
    points = points[:, list(use_dim)]
    if shift_height:
        floor_height = np.percentile(points[:, 2], 0.99)
        height = points[:, 2] - floor_height
        points = np.concatenate([points[:, :3], np.expand_dims(height, 1), points[:, 3:]], 1)
    return points


def load_points(path, load_dim=6, use_dim=(0, 1, 2), shift_height=True):
    """`.bin` (raw float32) or `.npy` file -> (N, C) float32 (loading.py:386-398)."""
    if str(path).endswith('.npy'):
        raw = np.load(path).astype(np.float32).reshape(-1)
    else:
        raw = np.fromfile(path, dtype=np.float32)
    return points_from_bytes(raw, load_dim, use_dim, shift_height)


def save_points(path, points):
    """Write (N, load_dim) float32 points in the reference's `.bin` layout."""
    np.ascontiguousarray(points, dtype=np.float32).tofile(path)


def indoor_point_sample(points, num_points, generator=None, choices=None):
    """IndoorPointSample on the device: points (N, C) tensor -> ((num_points, C), choices int64).
    Without replacement when N >= num_points, with replacement otherwise (transforms_3d.py:856-863).
    `choices` replays a given selection (the reference draws it with np.random.choice)."""
    n = points.shape[0]
    if choices is None:
        if n >= num_points:
            choices = torch.randperm(n, generator=generator, device=points.device)[:num_points]
        else:
            choices = torch.randint(0, n, (num_points,), generator=generator, device=points.device)
    choices = torch.as_tensor(choices, device=points.device, dtype=torch.long)
    return points.index_select(0, choices), choices


class SceneBatcher:
    """Files -> device batches.  `files`: list of `.bin` / `.npy` paths; yields (points (B, num_points,
    4) on `device`, list of per-scene choices).  The raw scenes are parsed once on the host (the
    percentile is numpy's, bit-identical to the reference), staged in pinned memory and sampled on
    the device."""

    def __init__(self, files, num_points=40000, batch_size=8, device='cuda', load_dim=6,
                 use_dim=(0, 1, 2), shift_height=True, seed=0):
        self.files, self.num_points, self.batch_size = list(files), num_points, batch_size
        self.device = torch.device(device)
        self.kw = dict(load_dim=load_dim, use_dim=use_dim, shift_height=shift_height)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self._cache = {}

    def _scene(self, i):
        if i not in self._cache:
            t = torch.from_numpy(load_points(self.files[i], **self.kw))
            self._cache[i] = t.pin_memory() if self.device.type == 'cuda' else t
        return self._cache[i]

    def batch(self, indices):
        pts, choices = [], []
        for i in indices:
            scene = self._scene(i).to(self.device, non_blocking=True)
            p, c = indoor_point_sample(scene, self.num_points, self.gen)
            pts.append(p)
            choices.append(c)
        return torch.stack(pts), choices

    def __iter__(self):
        for s in range(0, len(self.files) - self.batch_size + 1, self.batch_size):
            yield self.batch(range(s, s + self.batch_size))


def mean_teacher_views(points, gt_boxes=None, generator=None, rot_range=0.087266, scale_range=(1.0, 1.0),
                       trans_std=0.0):
    """One mean-teacher batch from sampled scenes: the student (strong) and teacher (weak) views of
    every scene with their augmentation records, and the GT boxes in the student frame.
    points (B, N, 4) on the device, gt_boxes (B, G, 7) padded or None."""
    B, dev = points.shape[0], points.device
    aug_s = BoxAug.random(B, dev, generator, rot_range, scale_range, trans_std)
    aug_t = BoxAug.random(B, dev, generator, rot_range, scale_range, trans_std)
    out = dict(points_s=aug_s.apply_points(points), points_t=aug_t.apply_points(points),
               aug_s=aug_s, aug_t=aug_t)
    if gt_boxes is not None:
        out['gt_boxes_s'] = transform_boxes(gt_boxes, aug_s)
    return out


def split_reference_state_dict(state):
    """A reference checkpoint's state_dict -> (module weights, ema_* buffers).  SimiTeacherHook keeps
    the EMA copies as model buffers `ema_<param name with '.' -> '_'>` (simi_teacher_hook.py:47-51), so
    `epoch_N.pth` / `epoch_N_ema.pth` carry them next to `backbone.*` / `bbox_head.*`."""
    weights = {k: v for k, v in state.items() if not k.startswith('ema_')}
    ema = {k: v for k, v in state.items() if k.startswith('ema_')}
    return weights, ema
