"""furthest_point_sample / furthest_point_sample_with_dist / Points_Sampler.

Host-side mirror of the reference's mmdet3d/ops/furthest_point_sample/
(furthest_point_sample.py:15-77, points_sampler.py:34-158, utils.py:4-31) on top of the
C ABI entry points nesie_fps / nesie_fps_with_dist.
"""
from typing import List

import torch
from torch import nn as nn
from torch.autograd import Function

from . import _lib


class FurthestPointSampling(Function):
    """D-FPS over xyz.  Same contract as the reference's Function of the same name
    (furthest_point_sample.py:15-35): (B, N, 3) contiguous fp32 -> (B, num_points) int32,
    non-differentiable, first sample is index 0."""

    @staticmethod
    def forward(ctx, points_xyz: torch.Tensor, num_points: int) -> torch.Tensor:
        assert points_xyz.is_contiguous()
        _lib.need_cuda(points_xyz)
        B, N = points_xyz.size()[:2]
        output = torch.empty((B, num_points), dtype=torch.int32, device=points_xyz.device)
        temp = None
        if _lib.lib().nesie_fps_needs_temp(B, N, num_points):
            temp = torch.full((B, N), 1e10, dtype=torch.float32, device=points_xyz.device)
        with torch.cuda.device(points_xyz.device):
            try:
                _lib.call("nesie_fps", B, N, num_points, _lib.ptr(points_xyz), _lib.ptr(temp),
                          _lib.ptr(output), _lib.stream())
            except RuntimeError:
                if temp is not None:
                    raise
                # the register-resident cluster kernel could not be scheduled (e.g. a cluster of 16 is
                # refused on this device / MIG slice): the global-memory kernel needs the reference's
                # `temp` buffer (furthest_point_sample.py:29-30)
                temp = torch.full((B, N), 1e10, dtype=torch.float32, device=points_xyz.device)
                _lib.call("nesie_fps", B, N, num_points, _lib.ptr(points_xyz), _lib.ptr(temp),
                          _lib.ptr(output), _lib.stream())
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


class FurthestPointSamplingWithDist(Function):
    """F-FPS over a precomputed (B, N, N) distance matrix (furthest_point_sample.py:42-74)."""

    @staticmethod
    def forward(ctx, points_dist: torch.Tensor, num_points: int) -> torch.Tensor:
        assert points_dist.is_contiguous()
        _lib.need_cuda(points_dist)
        B, N, _ = points_dist.size()
        output = points_dist.new_zeros([B, num_points], dtype=torch.int32)
        temp = points_dist.new_zeros([B, N]).fill_(1e10)
        with torch.cuda.device(points_dist.device):
            _lib.call("nesie_fps_with_dist", B, N, num_points, _lib.ptr(points_dist),
                      _lib.ptr(temp), _lib.ptr(output), _lib.stream())
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply
furthest_point_sample_with_dist = FurthestPointSamplingWithDist.apply


def calc_square_dist(point_feat_a, point_feat_b, norm=True):
    """Pairwise squared distance |a|^2 + |b|^2 - 2ab, (B,N,C) x (B,M,C) -> (B,N,M)
    (reference utils.py:4-31; optionally sqrt(d)/C)."""
    num_channel = point_feat_a.shape[-1]
    a_square = point_feat_a.pow(2).sum(dim=-1, keepdim=True)           # (B, N, 1)
    b_square = point_feat_b.pow(2).sum(dim=-1).unsqueeze(1)            # (B, 1, M)
    coor = torch.matmul(point_feat_a, point_feat_b.transpose(1, 2))
    dist = a_square + b_square - 2 * coor
    if norm:
        dist = torch.sqrt(dist) / num_channel
    return dist


class DFPS_Sampler(nn.Module):
    """FPS on Euclidean distance (points_sampler.py:104-116)."""

    def forward(self, points, features, npoint):
        return furthest_point_sample(points.contiguous(), npoint)


class FFPS_Sampler(nn.Module):
    """FPS on feature distance (points_sampler.py:119-135)."""

    def forward(self, points, features, npoint):
        assert features is not None, 'feature input to FFPS_Sampler should not be None'
        feats = torch.cat([points, features.transpose(1, 2)], dim=2)
        dist = calc_square_dist(feats, feats, norm=False)
        return furthest_point_sample_with_dist(dist.contiguous(), npoint)


class FS_Sampler(nn.Module):
    """F-FPS and D-FPS side by side (points_sampler.py:138-158)."""

    def forward(self, points, features, npoint):
        assert features is not None, 'feature input to FS_Sampler should not be None'
        feats = torch.cat([points, features.transpose(1, 2)], dim=2)
        dist = calc_square_dist(feats, feats, norm=False)
        idx_f = furthest_point_sample_with_dist(dist.contiguous(), npoint)
        idx_d = furthest_point_sample(points.contiguous(), npoint)
        return torch.cat([idx_f, idx_d], dim=1)


def get_sampler_type(sampler_type):
    samplers = {'D-FPS': DFPS_Sampler, 'F-FPS': FFPS_Sampler, 'FS': FS_Sampler}
    if sampler_type not in samplers:
        raise ValueError('Only "sampler_type" of "D-FPS", "F-FPS", or "FS"'
                         f' are supported, got {sampler_type}')
    return samplers[sampler_type]


class Points_Sampler(nn.Module):
    """Range-wise point sampling (points_sampler.py:34-101): sampler i is applied to
    points[last_end : fps_sample_range_list[i]] and its indices are offset by last_end."""

    def __init__(self, num_point: List[int], fps_mod_list: List[str] = ['D-FPS'],
                 fps_sample_range_list: List[int] = [-1]):
        super().__init__()
        assert len(num_point) == len(fps_mod_list) == len(fps_sample_range_list)
        self.num_point = num_point
        self.fps_sample_range_list = fps_sample_range_list
        self.samplers = nn.ModuleList(get_sampler_type(m)() for m in fps_mod_list)
        self.fp16_enabled = False

    def forward(self, points_xyz, features):
        if points_xyz.dtype != torch.float32:  # mmcv force_fp32
            points_xyz = points_xyz.float()
            features = features.float() if features is not None else None
        indices = []
        last_end = 0
        for rng, sampler, npoint in zip(self.fps_sample_range_list, self.samplers,
                                        self.num_point):
            assert rng < points_xyz.shape[1]
            if rng == -1:
                xyz_part = points_xyz[:, last_end:]
                feat_part = features[:, :, last_end:] if features is not None else None
            else:
                xyz_part = points_xyz[:, last_end:rng]
                feat_part = features[:, :, last_end:rng] if features is not None else None
            fps_idx = sampler(xyz_part.contiguous(), feat_part, npoint)
            indices.append(fps_idx + last_end)
            last_end += rng
        return torch.cat(indices, dim=1)
