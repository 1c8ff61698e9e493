"""three_nn / three_interpolate: mirror of mmdet3d/ops/interpolate/three_nn.py:8-45 and
three_interpolate.py:8-63."""
from typing import Tuple

import torch
from torch.autograd import Function

from . import _lib


class ThreeNN(Function):
    """For every target point its 3 nearest source points: (sqrt(d^2) (B,N,3) f32, idx int32)."""

    @staticmethod
    def forward(ctx, target: torch.Tensor, source: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert target.is_contiguous()
        assert source.is_contiguous()
        _lib.need_cuda(target, source)
        B, N, _ = target.size()
        m = source.size(1)
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=target.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=target.device)
        with torch.cuda.device(target.device):
            _lib.call("nesie_three_nn", B, N, m, _lib.ptr(target), _lib.ptr(source),
                      _lib.ptr(dist2), _lib.ptr(idx), _lib.stream())
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


def three_nn_grid(target, source, workspace=None):
    """three_nn through a uniform grid over the sources (nesie_three_nn_grid): bit-identical distances
    and indices, ~10x fewer distance evaluations when there are hundreds of sources or more.
    -> (sqrt(d^2) (B, N, 3), idx int32 (B, N, 3), workspace); pass the returned workspace back in to
    search further target sets against the SAME sources without binning them again."""
    assert target.is_contiguous() and source.is_contiguous()
    _lib.need_cuda(target, source)
    B, N, _ = target.size()
    m = source.size(1)
    dev = target.device
    dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((B, N, 3), dtype=torch.int32, device=dev)
    build = workspace is None
    if build:
        workspace = torch.empty((_lib.lib().nesie_ball_query_grid_workspace(B, m, 0),), dtype=torch.uint8,
                                device=dev)
    with torch.cuda.device(dev):
        _lib.call("nesie_three_nn_grid", B, N, m, _lib.ptr(target), _lib.ptr(source), _lib.ptr(dist2),
                  _lib.ptr(idx), _lib.ptr(workspace), workspace.numel(), int(build), _lib.stream())
    return torch.sqrt(dist2), idx, workspace


class ThreeInterpolate(Function):
    """out[b,c,j] = sum_i weight[b,j,i] * features[b,c,indices[b,j,i]]; differentiable w.r.t.
    features only (three_interpolate.py:11-60)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, indices: torch.Tensor,
                weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert indices.is_contiguous()
        assert weight.is_contiguous()
        _lib.need_cuda(features, indices, weight)
        B, c, m = features.size()
        n = indices.size(1)
        ctx.three_interpolate_for_backward = (indices, weight, m)
        output = torch.empty((B, c, n), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            _lib.call("nesie_three_interpolate", B, c, m, n, _lib.ptr(features),
                      _lib.ptr(indices), _lib.ptr(weight), _lib.ptr(output), _lib.stream())
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros((B, c, m), dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.data.contiguous()
        with torch.cuda.device(grad_out.device):
            _lib.call("nesie_three_interpolate_grad", B, c, n, m, _lib.ptr(grad_out_data),
                      _lib.ptr(idx), _lib.ptr(weight), _lib.ptr(grad_features), _lib.stream())
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply
