"""ball_query: host-side mirror of mmdet3d/ops/ball_query/ball_query.py:14-49 over nesie_ball_query."""
import os

import torch
from torch.autograd import Function

from . import _lib


GRID_MIN_PAIRS = 1 << 22  # centres x points above which the grid formulation is used


class BallQuery(Function):
    """First `sample_num` points (ascending index) with d2 == 0 or min_r^2 <= d2 < max_r^2 around
    every centre; short rows repeat their first hit, empty rows are zeros.  int32, non-diff."""

    @staticmethod
    def forward(ctx, min_radius: float, max_radius: float, sample_num: int, xyz: torch.Tensor,
                center_xyz: torch.Tensor) -> torch.Tensor:
        assert center_xyz.is_contiguous()
        assert xyz.is_contiguous()
        assert min_radius < max_radius
        _lib.need_cuda(xyz, center_xyz)
        B, N, _ = xyz.size()
        npoint = center_xyz.size(1)
        idx = torch.empty((B, npoint, sample_num), dtype=torch.int32, device=xyz.device)
        mode = os.environ.get("NESIE_BALL_QUERY", "auto")
        use_grid = mode == "grid" or (mode == "auto" and N * npoint >= GRID_MIN_PAIRS)
        use_grid = use_grid and N >= 1 and 0 < max_radius < 1e18 and sample_num <= 1024
        with torch.cuda.device(xyz.device):
            # reference launcher order: centres before points (ball_query.cpp:30-47)
            if use_grid:  # same output, via a uniform grid (large scenes)
                nbytes = _lib.lib().nesie_ball_query_grid_workspace(B, N, npoint)
                ws = torch.empty((nbytes,), dtype=torch.uint8, device=xyz.device)
                _lib.call("nesie_ball_query_grid", B, N, npoint, float(min_radius),
                          float(max_radius), sample_num, _lib.ptr(center_xyz), _lib.ptr(xyz),
                          _lib.ptr(idx), _lib.ptr(ws), nbytes, _lib.stream())
                _lib.LAUNCHES += 1  # build + query
            else:
                _lib.call("nesie_ball_query", B, N, npoint, float(min_radius), float(max_radius),
                          sample_num, _lib.ptr(center_xyz), _lib.ptr(xyz), _lib.ptr(idx),
                          _lib.stream())
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None, None


ball_query = BallQuery.apply
