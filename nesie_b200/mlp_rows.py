"""Shared MLP (1x1 conv -> BatchNorm(batch statistics) -> ReLU, repeated) on row-major activations
with the BatchNorm work fused into the tcgen05 GEMMs (training mode).

Same parameters and arithmetic as the ConvModule stacks of the reference's SA / FP modules
(ops/pointnet_modules/point_sa_module.py:136-158,279-288, point_fp_module.py:39-46).  Per layer the
unfused path sweeps HBM five times forward (GEMM read + write, statistics read, apply read + write);
here a layer is ONE kernel plus a tiny finalize:

  y_l = gemm( relu(bn_{l-1}(y_{l-1})) , W_l )   prologue: scale / shift + ReLU of the previous layer on
                                                the operand tile in shared memory (the activation is
                                                never written); epilogue: column sums of y_l and y_l^2
  stats_l = finalize(column sums)               mean / invstd / scale / shift + running statistics

The last layer's BatchNorm + ReLU (+ max-pool over the k rows of a group) runs as the fused
bn_relu_rows kernel fed with the GEMM's column sums.  Backward: weight gradient with the same
prologue on its X operand, data gradient, then the BatchNorm backward kernels (which only need y and
the saved statistics)."""
import os

import torch
from torch.autograd import Function

from . import _lib
from . import bn_rows
from .linear_rows import _pack, gemm_nt, sum_partials, wgrad


def enabled():
    return os.environ.get("NESIE_ROWS_FUSE", "1") != "0"


def _gemm_supported(x, n, k):
    return bool(_lib.lib().nesie_gemm_fused_supported(x.shape[0], n, k, _lib.ptr(x), k, n))


def supported(x, layers):
    """layers: [(weight (N, K), bn)]: every layer bias-free with a training-mode affine BatchNorm."""
    if not (enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and
            x.is_contiguous() and x.shape[0] >= 2):
        return False
    k = x.shape[1]
    if k > 512:     # the weight-gradient kernel of the backward pass takes k <= 512 (gemm_3xtf32.cu)
        return False
    for w, bn in layers:
        n = w.shape[0]
        if w.shape[1] != k or not (bn.training and bn.affine and bn.momentum is not None):
            return False
        if not (n % 4 == 0 and 4 <= n <= 256 and k % 4 == 0):
            return False
        k = n
    return _gemm_supported(x, layers[0][0].shape[0], x.shape[1])


def supported_eval(x, wa, wb):
    """Two chained GEMMs x @ wa.T -> (affine + ReLU prologue) @ wb.T on the fused kernel."""
    if not (enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and
            x.is_contiguous() and x.shape[0] >= 1 and wa.shape[1] == x.shape[1]):
        return False
    for w in (wa, wb):
        if not (w.shape[0] % 4 == 0 and 4 <= w.shape[0] <= 256 and w.shape[1] % 4 == 0):
            return False
    return wb.shape[1] == wa.shape[0] and _gemm_supported(x, wa.shape[0], x.shape[1])


def _gemm_fused(x, w, scale, shift, want_stats):
    """y = [relu(x * scale + shift)] @ w.T, optionally with the column-sum partials of y."""
    R, K = x.shape
    N = w.shape[0]
    y = torch.empty((R, N), dtype=torch.float32, device=x.device)
    parts = None
    if want_stats:
        nparts = _lib.lib().nesie_gemm_stats_parts(R)
        parts = torch.empty((nparts, 2, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        img = _pack(w, N, K, K, 1)
        _lib.call("nesie_gemm_nt_3xtf32_fused", R, N, K, _lib.ptr(x), K, _lib.ptr(img), _lib.ptr(y), N,
                  _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(parts), _lib.stream())
    return y, parts


def _wgrad_fused(gy, x, scale, shift):
    """gy^T @ relu(x * scale + shift) -> (N, K)."""
    R, N = gy.shape
    K = x.shape[1]
    ns = _lib.lib().nesie_gemm_wgrad_splits(R, N, K)
    parts = torch.empty((ns, N, K), dtype=torch.float32, device=gy.device)
    with torch.cuda.device(gy.device):
        _lib.call("nesie_gemm_wgrad_3xtf32_fused", R, N, K, _lib.ptr(gy), N, _lib.ptr(x), K,
                  _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(parts), ns, _lib.stream())
    return sum_partials(parts)


def bnbwd_fused_enabled():
    """Off by default: measured on the pretrain step (B200) the statistics in the data-gradient GEMM's
    epilogue cost more than the 0.9 ms sweep they replace (17.19 vs 16.23 ms per step, y tile read with
    coalesced 16-byte loads in the store layout; 17.65 with 4-byte loads): the epilogue of these GEMMs
    does not hide behind the next tile's MMAs, so every byte it touches is on the critical path.  Kept as
    an option (NESIE_BNBWD_FUSE=1) with its parity test."""
    return os.environ.get("NESIE_BNBWD_FUSE", "0") == "1"


def _dgrad_bn_backward(gy, w, y_prev, stats):
    """d_y_prev of  y = relu(bn(y_prev)) @ w.T : the data-gradient GEMM g = gy @ w takes the BatchNorm
    backward statistics (column sums of g * [relu active] and of that times xhat) in its epilogue, so
    the backward needs no statistics sweep over (g, y_prev): finalize + apply only.
    Returns (d_y_prev, d_gamma, d_beta)."""
    R, N = gy.shape
    C = y_prev.shape[1]
    dev = gy.device
    g_act = torch.empty((R, C), dtype=torch.float32, device=dev)
    nparts = _lib.lib().nesie_gemm_stats_parts(R)
    parts = torch.empty((nparts, 2, C), dtype=torch.float32, device=dev)
    d_y = torch.empty_like(y_prev)
    d_gamma = torch.empty((C,), dtype=torch.float32, device=dev)
    d_beta = torch.empty((C,), dtype=torch.float32, device=dev)
    ws = torch.empty((_lib.lib().nesie_bn_rows_workspace_bytes(C),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        img = _pack(w, C, N, 1, C)        # B = w^T (C x N): element (c, n) at w[n, c]
        _lib.call("nesie_gemm_nt_3xtf32_bnbwd", R, C, N, _lib.ptr(gy), N, _lib.ptr(img), _lib.ptr(g_act), C,
                  _lib.ptr(y_prev), C, _lib.ptr(stats), _lib.ptr(parts), _lib.stream())
        _lib.call("nesie_bn_relu_rows_backward_fused", R, C, _lib.ptr(y_prev), _lib.ptr(g_act),
                  _lib.ptr(stats), _lib.ptr(parts), nparts, _lib.ptr(d_y), _lib.ptr(d_gamma),
                  _lib.ptr(d_beta), _lib.ptr(ws), _lib.stream())
        _lib.LAUNCHES += 1
    return d_y, d_gamma, d_beta


def _bn_stats(y, parts, gamma, beta, rm, rv, eps, momentum):
    """[4, C] mean | invstd | scale | shift of y from the GEMM's column sums; updates rm / rv."""
    R, C = y.shape
    stats = torch.empty((4, C), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.call("nesie_bn_rows_forward_fused", R, C, 0, _lib.ptr(y), _lib.ptr(gamma), _lib.ptr(beta),
                  float(eps), float(momentum), _lib.ptr(rm), _lib.ptr(rv), _lib.ptr(parts),
                  parts.shape[0], _lib.ptr(stats), None, None, None, _lib.stream())
    return stats


def _fork2(f_w, f_d, like, shared):
    """Weight gradient beside the data-gradient chain of a layer's backward: the weight gradient only
    feeds the optimizer, the data gradient is what the rest of the backward waits for, and the two use
    different parts of an SM (tensor pipe + shared memory against HBM streaming in the BatchNorm kernels).
    NESIE_WGRAD_FORK=1 issues them as two forked branches (branches.py)."""
    if os.environ.get("NESIE_WGRAD_FORK", "0") != "1" or f_w is None or f_d is None:
        return (f_w() if f_w else None), (f_d() if f_d else None)
    from .branches import run_branches
    out = run_branches([f_w, f_d], like, shared, 2)
    return out[0], out[1]


class _LinearStats(Function):
    """First layer: y = x @ w.T plus the column sums of y."""

    @staticmethod
    def forward(ctx, x, w):
        y, parts = _gemm_fused(x, w, None, None, True)
        ctx.save_for_backward(x, w)
        ctx.mark_non_differentiable(parts)
        ctx.set_materialize_grads(False)  # no zero-filled gradient tensor for `parts`
        return y, parts

    @staticmethod
    def backward(ctx, gy, _gparts):
        x, w = ctx.saved_tensors
        if gy is None:
            return None, None
        gy = gy.contiguous()
        gx = gemm_nt(gy, w, transpose_w=True) if ctx.needs_input_grad[0] else None
        gw = wgrad(gy, x) if ctx.needs_input_grad[1] else None
        return gx, gw


class _BNReLULinear(Function):
    """y = relu(bn(y_prev)) @ w.T with bn's batch statistics taken from y_prev's column sums."""

    @staticmethod
    def forward(ctx, y_prev, parts_prev, gamma, beta, rm, rv, eps, momentum, w):
        stats = _bn_stats(y_prev, parts_prev, gamma, beta, rm, rv, eps, momentum)
        y, parts = _gemm_fused(y_prev, w, stats[2], stats[3], True)
        ctx.save_for_backward(y_prev, stats, w)
        ctx.mark_non_differentiable(parts)
        ctx.set_materialize_grads(False)
        return y, parts

    @staticmethod
    def backward(ctx, gy, _gparts):
        y_prev, stats, w = ctx.saved_tensors
        if gy is None:
            return (None,) * 9
        gy = gy.contiguous()
        R, C = y_prev.shape
        dev = y_prev.device
        if (bnbwd_fused_enabled() and C <= 256 and C % 4 == 0 and gy.shape[1] % 4 == 0 and
                _gemm_supported(gy, C, gy.shape[1])):
            gw = _wgrad_fused(gy, y_prev, stats[2], stats[3]) if ctx.needs_input_grad[8] else None
            d_y, d_gamma, d_beta = _dgrad_bn_backward(gy, w.contiguous(), y_prev, stats)
            return d_y, None, d_gamma, d_beta, None, None, None, None, gw

        def data_grad():
            g_act = gemm_nt(gy, w, transpose_w=True)          # gradient w.r.t. relu(bn(y_prev))
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

        f_w = (lambda: _wgrad_fused(gy, y_prev, stats[2], stats[3])) if ctx.needs_input_grad[8] else None
        gw, (d_y, d_gamma, d_beta) = _fork2(f_w, data_grad, gy, [gy, y_prev, stats, w])
        return d_y, None, d_gamma, d_beta, None, None, None, None, gw


def _pool_unit(k):
    """Rows per unit of the GEMM epilogue's pooling for groups of k rows (0: not available)."""
    if k == 16:
        return 16
    return 32 if (k % 32 == 0 and 32 <= k <= 224) else 0


def pooled_epilogue_enabled():
    return os.environ.get("NESIE_POOL_FUSE", "1") != "0"


class _BNReLULinearPooled(Function):
    """Last layer of a pooled shared MLP: y = relu(bn_prev(y_prev)) @ w.T, then relu(bn(y)) max-pooled over
    every k rows -- with the unit maxima / minima of y taken in the GEMM epilogue, so that y is written
    (the backward needs it) but not read again in the forward (nesie_bn_pool_finalize).
    Same gradients as _BNReLULinear followed by bn_rows._BNReLURows(k)."""

    @staticmethod
    def forward(ctx, y_prev, parts_prev, g0, b0, rm0, rv0, eps0, mom0, w, g1, b1, rm1, rv1, eps1, mom1, k):
        stats0 = _bn_stats(y_prev, parts_prev, g0, b0, rm0, rv0, eps0, mom0)
        R, K = y_prev.shape
        N = w.shape[0]
        dev = y_prev.device
        u = _pool_unit(k)
        y = torch.empty((R, N), dtype=torch.float32, device=dev)
        parts = torch.empty((_lib.lib().nesie_gemm_stats_parts(R), 2, N), dtype=torch.float32, device=dev)
        pmax = torch.empty((R // u, N), dtype=torch.float32, device=dev)
        pmin = torch.empty((R // u, N), dtype=torch.float32, device=dev)
        amax = torch.empty((R // u, N), dtype=torch.uint8, device=dev)
        amin = torch.empty((R // u, N), dtype=torch.uint8, device=dev)
        out = torch.empty((R // k, N), dtype=torch.float32, device=dev)
        arg = torch.empty((R // k, N), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            img = _pack(w, N, K, K, 1)
            _lib.call("nesie_gemm_nt_3xtf32_pool", R, N, K, _lib.ptr(y_prev), K, _lib.ptr(img), _lib.ptr(y), N,
                      _lib.ptr(stats0[2]), _lib.ptr(stats0[3]), _lib.ptr(parts), u, _lib.ptr(pmax),
                      _lib.ptr(amax), _lib.ptr(pmin), _lib.ptr(amin), None, 0, _lib.stream())
        stats1 = _bn_stats(y, parts, g1, b1, rm1, rv1, eps1, mom1)
        with torch.cuda.device(dev):
            _lib.call("nesie_bn_pool_finalize", R // k, k, u, N, _lib.ptr(pmax), _lib.ptr(amax), _lib.ptr(pmin),
                      _lib.ptr(amin), _lib.ptr(stats1), _lib.ptr(out), _lib.ptr(arg), _lib.stream())
        ctx.save_for_backward(y_prev, stats0, w, y, stats1, arg)
        ctx.k = k
        return out

    @staticmethod
    def backward(ctx, d_out):
        y_prev, stats0, w, y, stats1, arg = ctx.saved_tensors
        R, N = y.shape
        C = y_prev.shape[1]
        dev = y.device
        d_out = d_out.contiguous()
        # BatchNorm + ReLU + max-pool backward of this layer (bn_rows._BNReLURows.backward)
        gy = torch.empty_like(y)
        d_g1 = torch.empty((N,), dtype=torch.float32, device=dev)
        d_b1 = torch.empty((N,), dtype=torch.float32, device=dev)
        ws = torch.empty((_lib.lib().nesie_bn_rows_workspace_bytes(max(N, C)),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call("nesie_bn_relu_rows_backward", R, N, ctx.k, _lib.ptr(y), _lib.ptr(d_out), _lib.ptr(arg),
                      _lib.ptr(stats1), _lib.ptr(gy), _lib.ptr(d_g1), _lib.ptr(d_b1), _lib.ptr(ws), _lib.stream())
            _lib.LAUNCHES += 2
        # ... and of the GEMM with the previous layer's BatchNorm + ReLU in its prologue (_BNReLULinear.backward)
        gw = _wgrad_fused(gy, y_prev, stats0[2], stats0[3]) if ctx.needs_input_grad[8] else None
        g_act = gemm_nt(gy, w, transpose_w=True)
        d_y = torch.empty_like(y_prev)
        d_g0 = torch.empty((C,), dtype=torch.float32, device=dev)
        d_b0 = torch.empty((C,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("nesie_bn_relu_rows_backward", R, C, 0, _lib.ptr(y_prev), _lib.ptr(g_act), None,
                      _lib.ptr(stats0), _lib.ptr(d_y), _lib.ptr(d_g0), _lib.ptr(d_b0), _lib.ptr(ws), _lib.stream())
            _lib.LAUNCHES += 2
        return (d_y, None, d_g0, d_b0, None, None, None, None, gw, d_g1, d_b1, None, None, None, None, None)


def _bn_buffers(bn):
    if bn.track_running_stats:
        bn_rows.count_batch(bn)
        return bn.running_mean, bn.running_var
    return None, None


def mlp_rows(x, layers, pool_k=0):
    """x (R, K) -> relu(bn(... relu(bn(x @ W1^T)) ...)) as (R, C), or max-pooled over every pool_k
    consecutive rows -> (R / pool_k, C).  `supported(x, layers)` must hold."""
    w0, _ = layers[0]
    y, parts = _LinearStats.apply(x, w0)
    return mlp_rows_tail(y, parts, layers, pool_k)


def supported_tail(layers):
    """The layers after the first (whose output y (R, N1) and column sums come from elsewhere, e.g.
    gather_linear): bias-free convs with training-mode affine BatchNorms on the fused kernels."""
    if not enabled():
        return False
    k = layers[0][0].shape[0]
    if k > 512 or k % 4:
        return False
    for li, (w, bn) in enumerate(layers):
        n = w.shape[0]
        if not (bn.training and bn.affine and bn.momentum is not None):
            return False
        if li and not (w.shape[1] == k and n % 4 == 0 and 4 <= n <= 256):
            return False
        k = n
    return True


def mlp_rows_tail(y, parts, layers, pool_k=0):
    """mlp_rows from the first layer's pre-activation y = x @ W1^T and its column-sum partials on."""
    nl = len(layers)
    fuse_pool = (nl >= 2 and pool_k > 0 and pooled_epilogue_enabled() and _pool_unit(pool_k) > 0 and
                 y.shape[0] % pool_k == 0 and pool_k <= 254)
    for li in range(1, nl):
        if fuse_pool and li == nl - 1:
            bn0, bn1 = layers[li - 1][1], layers[li][1]
            rm0, rv0 = _bn_buffers(bn0)
            rm1, rv1 = _bn_buffers(bn1)
            return _BNReLULinearPooled.apply(y, parts, bn0.weight, bn0.bias, rm0, rv0, bn0.eps, bn0.momentum,
                                             layers[li][0], bn1.weight, bn1.bias, rm1, rv1, bn1.eps,
                                             bn1.momentum, pool_k)
        bn = layers[li - 1][1]
        rm, rv = _bn_buffers(bn)
        y, parts = _BNReLULinear.apply(y, parts, bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum,
                                       layers[li][0])
    return bn_rows.bn_relu_rows(y, layers[-1][1], pool_k, col_partials=parts)
