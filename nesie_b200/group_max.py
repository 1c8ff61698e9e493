"""Row-group max (+ bias, + broadcast-concat) for the SidePooling MiniPointNets: one kernel each way
instead of the reference formulation's max / expand / cat and their autograd kernels
(models/dense_heads/side_pooling_module.py:360-370)."""
import torch
from torch.autograd import Function

from . import _lib


class _GroupMaxRows(Function):

    @staticmethod
    def forward(ctx, x, bias, k, concat):
        x = x.contiguous()
        R, C = x.shape
        groups = R // k
        dev = x.device
        out = torch.empty((R, 2 * C) if concat else (groups, C), dtype=torch.float32, device=dev)
        arg = torch.empty((groups, C), dtype=torch.uint8, device=dev)
        bias_c = bias.contiguous() if bias is not None else None   # alive across the launch
        with torch.cuda.device(dev):
            _lib.call("nesie_group_max_rows_forward", groups, k, C, _lib.ptr(x),
                      _lib.ptr(bias_c), _lib.ptr(out),
                      _lib.ptr(arg), int(concat), _lib.stream())
        ctx.save_for_backward(arg)
        ctx.meta = (groups, k, C, bool(concat), bias is not None)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (arg,) = ctx.saved_tensors
        groups, k, C, concat, has_bias = ctx.meta
        d_out = d_out.contiguous()
        d_x = torch.empty((groups * k, C), dtype=torch.float32, device=d_out.device)
        want_bias = has_bias and ctx.needs_input_grad[1]
        nparts = _lib.lib().nesie_group_max_bias_parts(groups, C) if want_bias else 0
        parts = torch.empty((nparts, C), dtype=torch.float32, device=d_out.device) if nparts else None
        with torch.cuda.device(d_out.device):
            _lib.call("nesie_group_max_rows_backward", groups, k, C, _lib.ptr(d_out), _lib.ptr(arg),
                      _lib.ptr(d_x), int(concat), _lib.ptr(parts), _lib.stream())
        d_bias = None
        if want_bias:   # column sums of d_x: per-CTA partials from the kernel (a few thousand rows)
            d_bias = parts.sum(dim=0) if nparts else d_x.sum(dim=0)
        return d_x, d_bias, None, None


def supported(x, k):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] % 4 == 0 and
            1 <= k <= 255 and x.shape[0] % k == 0)


def group_max_rows(x, bias, k):
    """x (R, C) [+ bias (C)] -> (R / k, C): max over every k consecutive rows."""
    _lib.need_cuda(x)
    return _GroupMaxRows.apply(x, bias, k, False)


def group_max_concat_rows(x, bias, k):
    """x (R, C) [+ bias] -> (R, 2C): [ group max broadcast to the group's rows | x + bias ]."""
    _lib.need_cuda(x)
    return _GroupMaxRows.apply(x, bias, k, True)
