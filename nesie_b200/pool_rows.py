"""The MiniPointNet's two pooled stages (models/dense_heads/side_pooling_module.py:343-370) with the
max over a box's grid points taken in the row GEMM's epilogue (training mode).

    feature        = conv_b(relu(bn(conv_a(rows))))              (R, C)   first_conv
    feature_global = max over the k rows of a box                (R/k, C)
    feature        = second_conv(cat([feature_global.expand, feature]))
    out            = max over the k rows of a box                (R/k, F)

Reference formulation: max -> expand -> cat -> conv over 2C channels -> ... -> conv -> max, every
intermediate in HBM.  Here

  * `bn_relu_linear_max(..., store=True)`: conv_b's GEMM also emits the per-box maximum and its row
    from the accumulator tile (no second sweep over its output);
  * `concat_global_linear`: cat([g.expand, f]) @ W^T = f @ W_f^T + (g @ W_g^T)[box]: the global half is
    multiplied through the weights once per BOX and enters the GEMM over f as a per-group bias, so the
    (R, 2C) concatenation never exists and the contraction is half as long; backward: the group sums of
    the output gradient give the box term's gradient, the max's gradient is added in place;
  * `bn_relu_linear_max(..., store=False)`: the last conv's output is only ever max-pooled, so it is
    not written at all; its weight gradient is a gather of one input row per (box, channel)
    (`nesie_pool_wgrad`) instead of a contraction over a dense (R, F) gradient.
"""
import os

import torch
from torch.autograd import Function

from . import _lib
from . import mlp_rows
from .linear_rows import _pack, gemm_nt, sum_partials, wgrad


def enabled():
    return os.environ.get("NESIE_POOL_FUSE", "1") != "0"


def _sparse_dgrad():
    """nesie_pool_dgrad instead of scatter + dense GEMM for the pooled conv's data gradient.  Off by
    default: measured 87 us against 44 us (group_max backward + tcgen05 GEMM) at 4096 boxes x 16 rows,
    128 -> 256 channels -- the per-box sort and list walk are instruction-bound (profiles/r02_ncu_notes.md)."""
    return os.environ.get("NESIE_POOL_DGRAD", "0") == "1"


def pool_unit(k):
    """Rows per epilogue unit for groups of k rows, or 0 when the epilogue cannot pool them."""
    if k == 16:
        return 16
    return 32 if (k % 32 == 0 and 32 <= k <= 224) else 0


def _gemm_pool(x, img, N, scale, shift, want_stats, store, pool_k, grp_bias=None, grp_k=0):
    """x (R, K) [relu(x * scale + shift)] @ W^T with W's packed image `img` -> (y | None, column-sum
    partials | None, unit maxima | None, their rows | None)."""
    R, K = x.shape
    dev = x.device
    y = torch.empty((R, N), dtype=torch.float32, device=dev) if store else None
    parts = None
    if want_stats:
        parts = torch.empty((_lib.lib().nesie_gemm_stats_parts(R), 2, N), dtype=torch.float32, device=dev)
    u = pool_unit(pool_k) if pool_k else 0
    pmax = torch.empty((R // u, N), dtype=torch.float32, device=dev) if u else None
    amax = torch.empty((R // u, N), dtype=torch.uint8, device=dev) if u else None
    with torch.cuda.device(dev):
        _lib.call("nesie_gemm_nt_3xtf32_pool", R, N, K, _lib.ptr(x), K, _lib.ptr(img), _lib.ptr(y), N,
                  _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(parts), u, _lib.ptr(pmax), _lib.ptr(amax),
                  None, None, _lib.ptr(grp_bias), grp_k, _lib.stream())
    return y, parts, pmax, amax


def _finalize(pmax, amax, bias, k):
    u = pool_unit(k)
    U, N = pmax.shape
    G = U * u // k
    out = torch.empty((G, N), dtype=torch.float32, device=pmax.device)
    arg = torch.empty((G, N), dtype=torch.uint8, device=pmax.device)
    with torch.cuda.device(pmax.device):
        _lib.call("nesie_pool_finalize", G, k, u, N, _lib.ptr(pmax), _lib.ptr(amax), _lib.ptr(bias),
                  _lib.ptr(out), _lib.ptr(arg), _lib.stream())
    return out, arg


def _group_sum(x, k):
    G = x.shape[0] // k
    out = torch.empty((G, x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("nesie_group_sum_rows", G, k, x.shape[1], _lib.ptr(x), _lib.ptr(out), _lib.stream())
    return out


def _colsum(x):
    """Column sums of a (rows, n) matrix: one launch, fixed summation order (nesie_colsum_rows)."""
    rows, n = x.shape
    if n % 4 or n > 1024 or not x.is_contiguous():
        return x.sum(dim=0)
    out = torch.empty((n,), dtype=torch.float32, device=x.device)
    ws = torch.empty((_lib.lib().nesie_colsum_rows_workspace(n),), dtype=torch.uint8, device=x.device)
    ws[:256].zero_()                      # the kernel's arrival counter
    with torch.cuda.device(x.device):
        _lib.call("nesie_colsum_rows", rows, n, _lib.ptr(x), _lib.ptr(out), _lib.ptr(ws), _lib.stream())
    return out


class _AddBiasRows(Function):
    """x (R, n) + bias (n); the bias gradient is the one-launch column sum (ATen's reduction of a
    4096 x 128 matrix over dim 0 takes 13 us, x 21 Conv1d biases per SidePooling forward)."""

    @staticmethod
    def forward(ctx, x, bias):
        return x + bias

    @staticmethod
    def backward(ctx, gy):
        return gy, (_colsum(gy.contiguous()) if ctx.needs_input_grad[1] else None)


def add_bias_rows(x, bias):
    return _AddBiasRows.apply(x, bias) if x.is_cuda else x + bias


def _bn_relu_backward(y_prev, g_act, stats):
    """Gradient of relu(bn(y_prev)) w.r.t. y_prev / gamma / beta given g_act = dL/d(relu(bn(y_prev)))."""
    R, C = y_prev.shape
    dev = y_prev.device
    d_y = torch.empty_like(y_prev)
    d_gamma = torch.empty((C,), dtype=torch.float32, device=dev)
    d_beta = torch.empty((C,), dtype=torch.float32, device=dev)
    ws = torch.empty((_lib.lib().nesie_bn_rows_workspace_bytes(C),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("nesie_bn_relu_rows_backward", R, C, 0, _lib.ptr(y_prev), _lib.ptr(g_act), None,
                  _lib.ptr(stats), _lib.ptr(d_y), _lib.ptr(d_gamma), _lib.ptr(d_beta),
                  _lib.ptr(ws), _lib.stream())
        _lib.LAUNCHES += 2
    return d_y, d_gamma, d_beta


class _BNReLULinearMax(Function):
    """relu(bn(y_prev)) @ w.T with the maximum over every k rows (+ bias) from the epilogue.

    store=True : -> (y (R, N) WITHOUT the bias, max + bias (R/k, N), arg); only y carries a gradient
                 (the caller owns the bias and the maximum's gradient: concat_global_linear).
    store=False: -> max + bias (R/k, N); y is never written."""

    @staticmethod
    def forward(ctx, y_prev, parts_prev, gamma, beta, rm, rv, eps, momentum, w, bias, k, store):
        stats = mlp_rows._bn_stats(y_prev, parts_prev, gamma, beta, rm, rv, eps, momentum)
        N, K = w.shape
        with torch.cuda.device(y_prev.device):
            img = _pack(w, N, K, K, 1)
        y, _, pmax, amax = _gemm_pool(y_prev, img, N, stats[2], stats[3], False, store, k)
        bias_c = bias.contiguous() if bias is not None else None
        out, arg = _finalize(pmax, amax, bias_c, k)
        ctx.k, ctx.store, ctx.has_bias = k, store, bias is not None
        ctx.set_materialize_grads(False)
        if store:
            ctx.save_for_backward(y_prev, stats, w)
            ctx.mark_non_differentiable(out, arg)
            return y, out, arg
        ctx.save_for_backward(y_prev, stats, w, arg)
        return out

    @staticmethod
    def backward(ctx, *grads):
        none = (None,) * 12
        if ctx.store:
            y_prev, stats, w = ctx.saved_tensors
            gy = grads[0]
            if gy is None:
                return none
            gy = gy.contiguous()
            f_w = (lambda: mlp_rows._wgrad_fused(gy, y_prev, stats[2], stats[3])) if ctx.needs_input_grad[8] else None
            gw, (d_y, d_gamma, d_beta) = mlp_rows._fork2(
                f_w, lambda: _bn_relu_backward(y_prev, gemm_nt(gy, w, transpose_w=True), stats), gy,
                [gy, y_prev, stats, w])
            return d_y, None, d_gamma, d_beta, None, None, None, None, gw, None, None, None
        else:
            y_prev, stats, w, arg = ctx.saved_tensors
            d_out = grads[0]
            if d_out is None:
                return none
            d_out = d_out.contiguous()
            G, N = d_out.shape
            R, K = y_prev.shape
            dev = d_out.device
            gw = None
            with torch.cuda.device(dev):
                if ctx.needs_input_grad[8]:
                    parts = torch.empty((_lib.lib().nesie_pool_wgrad_parts(G), N, K), dtype=torch.float32,
                                        device=dev)
                    _lib.call("nesie_pool_wgrad", G, ctx.k, N, K, _lib.ptr(d_out), _lib.ptr(arg),
                              _lib.ptr(y_prev), _lib.ptr(stats[2]), _lib.ptr(stats[3]), _lib.ptr(parts),
                              _lib.stream())
                    gw = sum_partials(parts)
                if _sparse_dgrad() and ctx.k <= 64 and N * K <= 55296:   # weights fit its shared memory
                    w_c = w.contiguous()
                    g_act = torch.empty((R, K), dtype=torch.float32, device=dev)
                    _lib.call("nesie_pool_dgrad", G, ctx.k, N, K, _lib.ptr(d_out), _lib.ptr(arg),
                              _lib.ptr(w_c), _lib.ptr(g_act), _lib.stream())
                else:
                    gy = torch.empty((R, N), dtype=torch.float32, device=dev)
                    _lib.call("nesie_group_max_rows_backward", G, ctx.k, N, _lib.ptr(d_out), _lib.ptr(arg),
                              _lib.ptr(gy), 0, None, _lib.stream())
                    g_act = gemm_nt(gy, w, transpose_w=True)
            d_bias = _colsum(d_out) if (ctx.has_bias and ctx.needs_input_grad[9]) else None
        d_y, d_gamma, d_beta = _bn_relu_backward(y_prev, g_act, stats)
        return d_y, None, d_gamma, d_beta, None, None, None, None, gw, d_bias, None, None


class _ConcatGlobalLinear(Function):
    """cat([gmax.expand over the group's rows, y + bias]) @ w.T plus the column sums of the result;
    y (R, C) is conv_b's output WITHOUT its bias, gmax (R/k, C) = max over the group's rows of y + bias,
    arg its row, w (N, 2C) = [W_g | W_f].  Gradients: y (including the maximum's), bias, w."""

    @staticmethod
    def forward(ctx, y, gmax, arg, bias, w, k, zero_mean_grad):
        R, C = y.shape
        N = w.shape[0]
        dev = y.device
        with torch.cuda.device(dev):
            img_g = _pack(w, N, C, 2 * C, 1)              # columns 0 .. C-1 of w
            img_f = _pack(w[:, C:], N, C, 2 * C, 1)       # columns C .. 2C-1
            G = gmax.shape[0]
            e = torch.empty((G, N), dtype=torch.float32, device=dev)
            _lib.call("nesie_gemm_nt_3xtf32", G, N, C, _lib.ptr(gmax), C, _lib.ptr(img_g), _lib.ptr(e), N,
                      _lib.stream())
        e += torch.mv(w[:, C:], bias)                     # the bias of `y` through W_f: same for every row
        out, parts, _, _ = _gemm_pool(y, img_f, N, None, None, True, True, 0, grp_bias=e, grp_k=k)
        ctx.save_for_backward(y, gmax, arg, bias, w)
        ctx.k, ctx.zero_mean_grad = k, zero_mean_grad
        ctx.mark_non_differentiable(parts)
        ctx.set_materialize_grads(False)
        return out, parts

    @staticmethod
    def backward(ctx, g_out, _gparts):
        y, gmax, arg, bias, w = ctx.saved_tensors
        if g_out is None:
            return (None,) * 7
        g_out = g_out.contiguous()
        R, C = y.shape
        G = gmax.shape[0]
        dev = y.device
        w_g, w_f = w[:, :C].contiguous(), w[:, C:].contiguous()
        d_e = _group_sum(g_out, ctx.k)

        def data_grad():
            d_y = gemm_nt(g_out, w_f, transpose_w=True)           # (R, C)
            d_g = gemm_nt(d_e, w_g, transpose_w=True)             # (G, C) gradient of the group maximum
            with torch.cuda.device(dev):
                _lib.call("nesie_scatter_rows_add", G, ctx.k, C, _lib.ptr(d_g), _lib.ptr(arg), _lib.ptr(d_y),
                          _lib.stream())
            return d_y, d_g

        f_w = (lambda: (wgrad(d_e, gmax), wgrad(g_out, y))) if ctx.needs_input_grad[4] else None
        wg, (d_y, d_g) = mlp_rows._fork2(f_w, data_grad, g_out, [g_out, d_e, y, gmax, arg, w_g, w_f])
        # rows are y + bias: the bias reaches d_w_f and d_bias only through the column sums of g_out
        # (d_bias = colsum(g_out) @ (W_f + W_g)), which are exactly zero when g_out comes out of a
        # training-mode BatchNorm's backward -- a constant added in front of a BatchNorm has no
        # gradient; the reference formulation computes rounding noise there.  zero_mean_grad skips
        # those terms and returns d_bias = 0.
        d_w = d_bias = None
        if ctx.zero_mean_grad:
            if ctx.needs_input_grad[4]:
                d_w = torch.cat([wg[0], wg[1]], dim=1)
            if ctx.needs_input_grad[3]:
                d_bias = torch.zeros_like(bias)
        else:
            colsum = _colsum(d_e)
            if ctx.needs_input_grad[4]:
                d_w = torch.cat([wg[0], wg[1] + torch.outer(colsum, bias)], dim=1)
            if ctx.needs_input_grad[3]:
                d_bias = _colsum(d_g) + torch.mv(w_f.t(), colsum)     # sum over rows of d(y + bias)
        return d_y, None, None, d_bias, d_w, None, None


def bn_relu_linear_max(y_prev, parts_prev, bn, w, bias, k, store):
    rm, rv = mlp_rows._bn_buffers(bn)
    return _BNReLULinearMax.apply(y_prev, parts_prev, bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum,
                                  w, bias, k, store)


def concat_global_linear(y, gmax, arg, bias, w, k, zero_mean_grad=False):
    """zero_mean_grad: the result feeds a training-mode BatchNorm (its gradient has zero column sums)."""
    return _ConcatGlobalLinear.apply(y, gmax, arg, bias, w, k, zero_mean_grad)
