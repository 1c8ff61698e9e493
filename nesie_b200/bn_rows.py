"""Training-mode BatchNorm + ReLU (+ max-pool over the nsample rows of a group) on row-major
activations, as one fused pair of kernels each way (nesie_bn_relu_rows_forward / _backward).

Same parameters, statistics and running-stat updates as the nn.BatchNorm2d + ReLU + F.max_pool2d
tail of the reference's ConvModules (ops/pointnet_modules/point_sa_module.py:149-150,279-288)."""
import torch
from torch.autograd import Function

from . import _lib


_deferred = None   # list of counters while inside defer_batch_counters()


class defer_batch_counters:
    """Inside this context the `num_batches_tracked += 1` of every BatchNorm evaluated by this repo's
    row kernels is collected and applied as ONE multi-tensor launch on exit (23 one-element kernels
    per training step otherwise); the counters end up exactly as nn.BatchNorm leaves them."""

    def __enter__(self):
        global _deferred
        self._outer = _deferred
        _deferred = []
        return self

    def __exit__(self, *exc):
        global _deferred
        pending, _deferred = _deferred, self._outer
        if pending:
            torch._foreach_add_(pending, 1)
        return False


def count_batch(bn):
    """bn.num_batches_tracked += 1 (deferred inside defer_batch_counters())."""
    if _deferred is not None:
        _deferred.append(bn.num_batches_tracked)
    else:
        bn.num_batches_tracked.add_(1)


def supported(y, bn, k=0):
    R, C = y.shape
    return (y.is_cuda and y.dtype == torch.float32 and bn.training and bn.affine and
            bn.momentum is not None and C % 4 == 0 and 4 <= C <= 1024 and R >= 2 and
            0 <= k <= 254 and (k == 0 or R % k == 0))


class _BNReLURows(Function):

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, eps, momentum, k, col_partials=None):
        y = y.contiguous()
        R, C = y.shape
        dev = y.device
        stats = torch.empty((4, C), dtype=torch.float32, device=dev)
        ws = torch.empty((_lib.lib().nesie_bn_rows_workspace_bytes(C),), dtype=torch.uint8, device=dev)
        if k == 0:
            out = torch.empty((R, C), dtype=torch.float32, device=dev)
            arg = None
        else:
            out = torch.empty((R // k, C), dtype=torch.float32, device=dev)
            arg = torch.empty((R // k, C), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            if col_partials is None:
                _lib.call("nesie_bn_relu_rows_forward", R, C, k, _lib.ptr(y), _lib.ptr(gamma),
                          _lib.ptr(beta), float(eps), float(momentum), _lib.ptr(running_mean),
                          _lib.ptr(running_var), _lib.ptr(stats), _lib.ptr(out), _lib.ptr(arg),
                          _lib.ptr(ws), _lib.stream())
                _lib.LAUNCHES += 2
            else:  # statistics from the producing GEMM's column sums: no sweep over y for them
                _lib.call("nesie_bn_rows_forward_fused", R, C, k, _lib.ptr(y), _lib.ptr(gamma),
                          _lib.ptr(beta), float(eps), float(momentum), _lib.ptr(running_mean),
                          _lib.ptr(running_var), _lib.ptr(col_partials), col_partials.shape[0],
                          _lib.ptr(stats), _lib.ptr(out), _lib.ptr(arg), _lib.ptr(ws), _lib.stream())
                _lib.LAUNCHES += 1
        ctx.save_for_backward(y, stats, arg)
        ctx.k = k
        ctx.mark_non_differentiable(running_mean, running_var) if False else None
        return out

    @staticmethod
    def backward(ctx, d_out):
        y, stats, arg = ctx.saved_tensors
        R, C = y.shape
        dev = y.device
        d_out = d_out.contiguous()
        d_y = torch.empty_like(y)
        d_gamma = torch.empty((C,), dtype=torch.float32, device=dev)
        d_beta = torch.empty((C,), dtype=torch.float32, device=dev)
        ws = torch.empty((_lib.lib().nesie_bn_rows_workspace_bytes(C),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call("nesie_bn_relu_rows_backward", R, C, ctx.k, _lib.ptr(y), _lib.ptr(d_out),
                      _lib.ptr(arg), _lib.ptr(stats), _lib.ptr(d_y), _lib.ptr(d_gamma),
                      _lib.ptr(d_beta), _lib.ptr(ws), _lib.stream())
            _lib.LAUNCHES += 2
        return d_y, d_gamma, d_beta, None, None, None, None, None, None


def bn_relu_rows(y, bn, k=0, col_partials=None):
    """relu(batch_norm(y)) for y (R, C) with `bn`'s parameters in training mode; k > 0 additionally
    max-pools every k consecutive rows -> (R/k, C).  Updates bn's running statistics.
    col_partials: optional (nparts, 2, C) column sums of y and y^2 from the producing GEMM."""
    if bn.track_running_stats:
        count_batch(bn)
        rm, rv = bn.running_mean, bn.running_var
    else:
        rm = rv = None
    return _BNReLURows.apply(y, bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum, k, col_partials)
