"""y = sum_j w_j * table[idx_j] + head @ wx^T: a linear layer commuted with the gather in front of it
(nesie_gather_linear_forward / _backward, csrc/gather_linear.cu).  `table` is the layer's weight
already applied to the few thousand source rows; gradients flow to `table` and `wx`."""
import torch
from torch.autograd import Function

from . import _lib
from .linear_rows import sum_partials


def supported(c):
    t = c // 4
    return c % 4 == 0 and 1 <= t <= 256 and (t & (t - 1)) == 0


class _GatherLinear(Function):

    @staticmethod
    def forward(ctx, table, idx, weight, head, wx, want_stats, xyz=None, center=None, ns=1, radius=0.0):
        _lib.need_cuda(table, idx, weight, head, wx, xyz, center)
        table = table.contiguous()
        B, M, C = table.shape
        n, J = idx.shape[1], idx.shape[2]
        dev = table.device
        y = torch.empty((B * n, C), dtype=torch.float32, device=dev)
        parts = None
        if want_stats:
            parts = torch.empty((_lib.lib().nesie_gather_linear_parts(B, C, n), 2, C), dtype=torch.float32,
                                device=dev)
        wx_c = wx.contiguous() if wx is not None else None
        with torch.cuda.device(dev):
            _lib.call("nesie_gather_linear_forward", B, C, M, n, J, _lib.ptr(table), _lib.ptr(idx),
                      _lib.ptr(weight), _lib.ptr(head), _lib.ptr(wx_c), _lib.ptr(xyz), _lib.ptr(center),
                      ns, float(radius), _lib.ptr(y), _lib.ptr(parts), _lib.stream())
        ctx.save_for_backward(idx, weight, head, xyz, center)
        ctx.shape = (B, M, C, n, J)
        ctx.ns, ctx.radius = ns, float(radius)
        ctx.has_wx = wx is not None
        if parts is not None:
            ctx.mark_non_differentiable(parts)
        ctx.set_materialize_grads(False)
        return y, parts

    @staticmethod
    def backward(ctx, d_y, _gparts):
        idx, weight, head, xyz, center = ctx.saved_tensors
        if d_y is None:
            return (None,) * 10
        B, M, C, n, J = ctx.shape
        d_y = d_y.contiguous()
        dev = d_y.device
        d_table = torch.zeros((B, M, C), dtype=torch.float32, device=dev)
        want_wx = ctx.has_wx and ctx.needs_input_grad[4]
        parts = None
        if want_wx:
            parts = torch.empty((_lib.lib().nesie_gather_linear_parts(B, C, n), C, 4), dtype=torch.float32,
                                device=dev)
        with torch.cuda.device(dev):
            _lib.call("nesie_gather_linear_backward", B, C, M, n, J, _lib.ptr(d_y), _lib.ptr(idx),
                      _lib.ptr(weight), _lib.ptr(head if want_wx else None),
                      _lib.ptr(xyz if want_wx else None), _lib.ptr(center if want_wx else None), ctx.ns,
                      ctx.radius, _lib.ptr(d_table), _lib.ptr(parts), _lib.stream())
        d_wx = sum_partials(parts)[:, :3] if want_wx else None
        return ((d_table if ctx.needs_input_grad[0] else None), None, None, None, d_wx, None, None, None,
                None, None)


def gather_linear(table, idx, weight=None, head=None, wx=None, want_stats=False, xyz=None, center=None,
                  ns=1, radius=0.0):
    """table (B, M, C), idx (B, n, J) int32 [, weight (B, n, J)], wx (C, 3) with head (B, n, 3) or with
    xyz (B, M, 3) + center (B, n / ns, 3) (head = (xyz[idx] - center) / radius, no gradient to either) ->
    (y (B * n, C), column-sum partials (parts, 2, C) | None)."""
    return _GatherLinear.apply(table, idx, weight, head, wx, want_stats, xyz, center, ns, radius)
