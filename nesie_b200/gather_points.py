"""gather_points: mirror of mmdet3d/ops/gather_points/gather_points.py:7-52."""
import torch
from torch.autograd import Function

from . import _lib


class GatherPoints(Function):
    """out[b, c, j] = features[b, c, indices[b, j]]; differentiable w.r.t. features."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert indices.is_contiguous()
        _lib.need_cuda(features, indices)
        B, npoint = indices.size()
        _, C, N = features.size()
        output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            _lib.call("nesie_gather_points", B, C, N, npoint, _lib.ptr(features),
                      _lib.ptr(indices), _lib.ptr(output), _lib.stream())
        ctx.for_backwards = (indices, C, N)
        ctx.mark_non_differentiable(indices)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.data.contiguous()
        with torch.cuda.device(grad_out.device):
            _lib.call("nesie_gather_points_grad", B, C, N, npoint, _lib.ptr(grad_out_data),
                      _lib.ptr(idx), _lib.ptr(grad_features), _lib.stream())
        return grad_features, None


gather_points = GatherPoints.apply
