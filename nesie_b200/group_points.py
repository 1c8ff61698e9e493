"""grouping_operation / QueryAndGroup / GroupAll.

Mirror of mmdet3d/ops/group_points/group_points.py:12-226.  `grouping_operation` is the plain
drop-in op; `QueryAndGroup` keeps the reference's constructor and outputs but runs its body
(group xyz, subtract centre, divide by radius, group features, concat) as ONE kernel,
nesie_query_group_concat, instead of five launches.
"""
from typing import Tuple

import torch
from torch import nn as nn
from torch.autograd import Function

from . import _lib
from .ball_query import ball_query


class GroupingOperation(Function):
    """out[b, c, j, k] = features[b, c, indices[b, j, k]] (group_points.py:172-224)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert indices.is_contiguous()
        _lib.need_cuda(features, indices)
        B, nfeatures, nsample = indices.size()
        _, C, N = features.size()
        output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32,
                             device=features.device)
        with torch.cuda.device(features.device):
            _lib.call("nesie_group_points", B, C, N, nfeatures, nsample, _lib.ptr(features),
                      _lib.ptr(indices), _lib.ptr(output), _lib.stream())
        ctx.for_backwards = (indices, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.data.contiguous()
        with torch.cuda.device(grad_out.device):
            _lib.call("nesie_group_points_grad", B, C, N, npoint, nsample,
                      _lib.ptr(grad_out_data), _lib.ptr(idx), _lib.ptr(grad_features),
                      _lib.stream())
        return grad_features, None


grouping_operation = GroupingOperation.apply


class _QueryGroupConcat(Function):
    """(xyz[idx] - centre) * (1/radius)  ++  features[idx]  ->  (B, 3 + C, npoint, nsample).

    Gradients follow the unfused reference graph (group_points.py:98-116): to `features` through
    the feature gather, to `points_xyz` through the xyz gather and to `center_xyz` through the
    subtraction."""

    @staticmethod
    def forward(ctx, points_xyz, center_xyz, features, idx, radius):
        _lib.need_cuda(points_xyz, center_xyz, features, idx)
        points_xyz = points_xyz.contiguous()
        center_xyz = center_xyz.contiguous()
        B, N, _ = points_xyz.shape
        npoint, nsample = idx.shape[1], idx.shape[2]
        C = 0
        if features is not None:
            features = features.contiguous()
            C = features.shape[1]
        out = torch.empty((B, 3 + C, npoint, nsample), dtype=torch.float32,
                          device=points_xyz.device)
        with torch.cuda.device(points_xyz.device):
            _lib.call("nesie_query_group_concat", B, C, N, npoint, nsample, _lib.ptr(points_xyz),
                      _lib.ptr(center_xyz), _lib.ptr(features), _lib.ptr(idx), float(radius),
                      _lib.ptr(out), _lib.stream())
        ctx.save_for_backward(idx)
        ctx.shape = (B, C, N, npoint, nsample)
        ctx.radius = float(radius)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, N, npoint, nsample = ctx.shape
        g_xyz = g_center = g_feat = None
        dev = grad_out.device
        with torch.cuda.device(dev):
            if ctx.needs_input_grad[2] and C > 0:
                g = grad_out[:, 3:].contiguous()
                g_feat = torch.zeros((B, C, N), dtype=torch.float32, device=dev)
                _lib.call("nesie_group_points_grad", B, C, N, npoint, nsample, _lib.ptr(g),
                          _lib.ptr(idx), _lib.ptr(g_feat), _lib.stream())
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                g = grad_out[:, :3]
                if ctx.radius > 0:
                    g = g * (torch.tensor(1.0, dtype=torch.float32) /
                             torch.tensor(ctx.radius, dtype=torch.float32)).item()
                g = g.contiguous()
                if ctx.needs_input_grad[0]:
                    gt = torch.zeros((B, 3, N), dtype=torch.float32, device=dev)
                    _lib.call("nesie_group_points_grad", B, 3, N, npoint, nsample, _lib.ptr(g),
                              _lib.ptr(idx), _lib.ptr(gt), _lib.stream())
                    g_xyz = gt.transpose(1, 2).contiguous()
                if ctx.needs_input_grad[1]:
                    g_center = -g.sum(dim=3).transpose(1, 2).contiguous()
        return g_xyz, g_center, g_feat, None, None


class _QueryGroupRows(Function):
    """Row-major grouped tensor (B*npoint*nsample, 3 + C) for the GEMM formulation of the shared
    MLP: [ (xyz[idx] - centre) * (1/radius) | features[idx] ] per row.  Same values as
    _QueryGroupConcat, different layout; gradients to features, points_xyz and center_xyz.
    pad_to > 1 rounds the row width up to a multiple of it with zero columns (16-byte aligned rows
    for the GEMM's vector loads; the matching weight columns are zero-padded by the caller)."""

    @staticmethod
    def forward(ctx, points_xyz, center_xyz, features, idx, radius, pad_to=1):
        _lib.need_cuda(points_xyz, center_xyz, features, idx)
        points_xyz = points_xyz.contiguous()
        center_xyz = center_xyz.contiguous()
        B, N, _ = points_xyz.shape
        npoint, nsample = idx.shape[1], idx.shape[2]
        C = 0 if features is None else features.shape[1]
        table = None if features is None else features.transpose(1, 2).contiguous()  # (B, N, C)
        ld = -(-(3 + C) // pad_to) * pad_to
        rows = torch.empty((B * npoint * nsample, ld), dtype=torch.float32,
                           device=points_xyz.device)
        with torch.cuda.device(points_xyz.device):
            _lib.call("nesie_group_rows", B, C, N, npoint, nsample, _lib.ptr(points_xyz),
                      _lib.ptr(center_xyz), _lib.ptr(table), _lib.ptr(idx), float(radius),
                      _lib.ptr(rows), ld, _lib.stream())
        ctx.save_for_backward(idx)
        ctx.shape = (B, C, N, npoint, nsample)
        ctx.ld = ld
        ctx.radius = float(radius)
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        (idx,) = ctx.saved_tensors
        B, C, N, npoint, nsample = ctx.shape
        dev = grad_rows.device
        grad_rows = grad_rows.contiguous()
        g_table = torch.zeros((B, N, C), dtype=torch.float32, device=dev) \
            if (ctx.needs_input_grad[2] and C > 0) else None
        g_xyz = torch.zeros((B, N, 3), dtype=torch.float32, device=dev) \
            if ctx.needs_input_grad[0] else None
        g_center = torch.zeros((B, npoint, 3), dtype=torch.float32, device=dev) \
            if ctx.needs_input_grad[1] else None
        if g_table is not None or g_xyz is not None or g_center is not None:
            with torch.cuda.device(dev):
                _lib.call("nesie_group_rows_grad", B, C, N, npoint, nsample, _lib.ptr(grad_rows),
                          _lib.ptr(idx), ctx.radius, _lib.ptr(g_table), _lib.ptr(g_xyz),
                          _lib.ptr(g_center), ctx.ld, _lib.stream())
        g_feat = g_table.transpose(1, 2) if g_table is not None else None  # view: no copy
        return g_xyz, g_center, g_feat, None, None, None


class QueryAndGroup(nn.Module):
    """Ball query + grouping.  Constructor arguments, assertions and return values as in the
    reference (group_points.py:36-128).  kNN grouping (max_radius=None) and uniform_sample are
    not on the Nesie/VoteNet path (every SA layer passes a radius) and are not implemented."""

    def __init__(self, max_radius, sample_num, min_radius=0, use_xyz=True,
                 return_grouped_xyz=False, normalize_xyz=False, uniform_sample=False,
                 return_unique_cnt=False, return_grouped_idx=False):
        super().__init__()
        self.max_radius = max_radius
        self.min_radius = min_radius
        self.sample_num = sample_num
        self.use_xyz = use_xyz
        self.return_grouped_xyz = return_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.uniform_sample = uniform_sample
        self.return_unique_cnt = return_unique_cnt
        self.return_grouped_idx = return_grouped_idx
        if self.return_unique_cnt:
            assert self.uniform_sample, \
                'uniform_sample should be True when returning the count of unique samples'
        if self.max_radius is None:
            assert not self.normalize_xyz, \
                'can not normalize grouped xyz when max_radius is None'
            raise NotImplementedError('kNN grouping (max_radius=None) is outside the Nesie hot path')
        if self.uniform_sample:
            raise NotImplementedError('uniform_sample is outside the Nesie hot path')

    def forward_rows(self, points_xyz, center_xyz, features, pad_to=1):
        """Row-major variant for the GEMM-form shared MLP: (B*npoint*sample_num, 3+C) rows
        (row width rounded up to a multiple of pad_to with zero columns)."""
        idx = ball_query(self.min_radius, self.max_radius, self.sample_num,
                         points_xyz.contiguous(), center_xyz.contiguous())
        radius = self.max_radius if self.normalize_xyz else 0.0
        return _QueryGroupRows.apply(points_xyz, center_xyz, features, idx, radius, pad_to)

    def forward(self, points_xyz, center_xyz, features=None):
        """points_xyz (B,N,3), center_xyz (B,npoint,3), features (B,C,N) ->
        (B, 3+C, npoint, sample_num) [, grouped_xyz] [, idx]."""
        idx = ball_query(self.min_radius, self.max_radius, self.sample_num,
                         points_xyz.contiguous(), center_xyz.contiguous())
        radius = self.max_radius if self.normalize_xyz else 0.0
        if features is None:
            assert self.use_xyz, 'Cannot have not features and not use xyz as a feature!'
        grouped = _QueryGroupConcat.apply(points_xyz, center_xyz, features, idx, radius)
        grouped_xyz = grouped[:, :3]
        if features is not None and not self.use_xyz:
            new_features = grouped[:, 3:]
        else:
            new_features = grouped
        ret = [new_features]
        if self.return_grouped_xyz:
            ret.append(grouped_xyz)
        if self.return_grouped_idx:
            ret.append(idx)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """Group every point into one region (group_points.py:131-169)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)
        return grouped_features
