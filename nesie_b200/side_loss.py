"""Per-side uncertainty box-regression loss as one fused kernel each way.

Mirror of NesieHead.loss lines nesie_head.py:332-349 + SurfaceLoss (MSE branch,
surface_loss.py:57-61) + Bbox2Surface (surface_loss.py:90-100) of the reference."""
import torch
from torch.autograd import Function

from . import _lib


def bbox2surface(bbox):
    """(..., >=6) [cx,cy,cz,sx,sy,sz,..] -> (..., 6) [min xyz, max xyz] (surface_loss.py:90-100)."""
    center, size = bbox[..., :3], bbox[..., 3:6]
    return torch.cat([center - 0.5 * size, center + 0.5 * size], dim=-1)


class _SideUncertaintyLoss(Function):

    @staticmethod
    def forward(ctx, surface_pred, box_targets, side_scores, sem_scores, weight, loss_weight,
                alpha):
        _lib.need_cuda(surface_pred, box_targets, side_scores, sem_scores, weight)
        surface_pred = surface_pred.contiguous().float()
        box_targets = box_targets.contiguous().float()
        side_scores = side_scores.contiguous().float()
        sem_scores = sem_scores.contiguous().float()
        weight = weight.contiguous().float()
        rows, ncls = sem_scores.shape
        assert surface_pred.shape == (rows, 6) and box_targets.shape[0] == rows
        assert box_targets.shape[1] == 7 and side_scores.shape == (rows, 6, ncls)
        loss = torch.zeros((), dtype=torch.float32, device=surface_pred.device)
        sigma = torch.empty((rows, 6), dtype=torch.float32, device=surface_pred.device)
        with torch.cuda.device(surface_pred.device):
            _lib.call("nesie_side_uncertainty_loss", rows, ncls, _lib.ptr(surface_pred),
                      _lib.ptr(box_targets), _lib.ptr(side_scores), _lib.ptr(sem_scores),
                      _lib.ptr(weight), float(loss_weight), float(alpha), _lib.ptr(loss),
                      _lib.ptr(sigma), _lib.stream())
        ctx.save_for_backward(surface_pred, box_targets, side_scores, sem_scores, weight)
        ctx.cfg = (float(loss_weight), float(alpha))
        return loss, sigma

    @staticmethod
    def backward(ctx, grad_loss, grad_sigma):
        surface_pred, box_targets, side_scores, sem_scores, weight = ctx.saved_tensors
        loss_weight, alpha = ctx.cfg
        rows, ncls = sem_scores.shape
        dev = surface_pred.device
        g_pred = torch.empty_like(surface_pred) if ctx.needs_input_grad[0] else None
        g_side = torch.zeros_like(side_scores) if ctx.needs_input_grad[2] else None
        grad_loss = grad_loss.contiguous().float().reshape(1)
        gsig = grad_sigma.contiguous().float() if grad_sigma is not None else None
        with torch.cuda.device(dev):
            _lib.call("nesie_side_uncertainty_loss_grad", rows, ncls, _lib.ptr(surface_pred),
                      _lib.ptr(box_targets), _lib.ptr(side_scores), _lib.ptr(sem_scores),
                      _lib.ptr(weight), loss_weight, alpha, _lib.ptr(grad_loss), _lib.ptr(gsig),
                      _lib.ptr(g_pred), _lib.ptr(g_side), _lib.stream())
        return g_pred, None, g_side, None, None, None, None


def side_uncertainty_loss(surface_pred, box_targets, side_scores, sem_scores, weight,
                          loss_weight=10.0, alpha=1.0):
    """surface_pred (rows,6), box_targets (rows,7), side_scores (rows,6,C), sem_scores (rows,C),
    weight (rows,6) -> (loss scalar, sigma (rows,6)).  loss = sum(exp(-sigma) * L + alpha * sigma
    * w) with L = loss_weight * w * (pred - Bbox2Surface(target))^2 and sigma the side-score
    polynomial 0.8 s^2 - 1.8 s + 1 at the argmax class.  Gradients flow to surface_pred and
    side_scores (also through the returned sigma)."""
    return _SideUncertaintyLoss.apply(surface_pred, box_targets, side_scores, sem_scores, weight,
                                      loss_weight, alpha)
