"""Row-major linear layer on the tcgen05 3xTF32 GEMM (fp32 parity), with autograd.

y = x @ w.T for x (R, K) and w (N, K): the training-mode 1x1 convolution of the SA / FP shared
MLPs in its GEMM form.  Forward and the data gradient run on nesie_gemm_nt_3xtf32; the weight
gradient (a reduction over the R rows) is a library GEMM."""
import torch
from torch.autograd import Function

from . import _lib


def _pack(w, n, k, stride_n, stride_k):
    nbytes = _lib.lib().nesie_gemm_b_image_bytes(n, k)
    img = torch.empty((nbytes,), dtype=torch.uint8, device=w.device)
    _lib.call("nesie_gemm_pack_b", n, k, stride_n, stride_k, _lib.ptr(w), _lib.ptr(img),
              _lib.stream())
    return img


def gemm_nt(a, w, transpose_w=False):
    """a (R, K) fp32 contiguous; w (N, K) [or (K, N) with transpose_w] -> (R, N) fp32."""
    _lib.need_cuda(a, w)
    a = a.contiguous()
    w = w.contiguous()
    R, K = a.shape
    if transpose_w:
        assert w.shape[0] == K
        N = w.shape[1]
        sn, sk = 1, N
    else:
        assert w.shape[1] == K
        N = w.shape[0]
        sn, sk = K, 1
    out = torch.empty((R, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        # the kernel produces at most 256 output columns (one TMEM accumulator): wider products
        # (the data gradient of a layer with more than 256 input channels) run as column blocks
        for n0 in range(0, N, 256):
            nb = min(256, N - n0)
            img = _pack(w.reshape(-1)[n0 * sn:], nb, K, sn, sk)
            _lib.call("nesie_gemm_nt_3xtf32", R, nb, K, _lib.ptr(a), K, _lib.ptr(img),
                      _lib.ptr(out) + 4 * n0, N, _lib.stream())
    return out


def wgrad(gy, x):
    """gy (R, N), x (R, K) fp32 contiguous -> gy^T @ x (N, K) on the tensor cores (3xTF32)."""
    _lib.need_cuda(gy, x)
    R, N = gy.shape
    K = x.shape[1]
    if N > 256:     # the kernel takes at most 256 rows of the result: column blocks of gy (row stride N)
        return torch.cat([_wgrad_block(gy, x, n0, min(256, N - n0)) for n0 in range(0, N, 256)], dim=0)
    return _wgrad_block(gy, x, 0, N)


def _wgrad_block(gy, x, n0, nb):
    R, N = gy.shape
    K = x.shape[1]
    ns = _lib.lib().nesie_gemm_wgrad_splits(R, nb, K)
    parts = torch.empty((ns, nb, K), dtype=torch.float32, device=gy.device)
    with torch.cuda.device(gy.device):
        _lib.call("nesie_gemm_wgrad_3xtf32", R, nb, K, _lib.ptr(gy) + 4 * n0, N, _lib.ptr(x), K,
                  _lib.ptr(parts), ns, _lib.stream())
    return sum_partials(parts)


def sum_partials(parts):
    """(ns, N, K) partial blocks -> (N, K), added in ascending order (deterministic)."""
    ns, N, K = parts.shape
    if ns == 1:
        return parts[0]
    if (N * K) % 4:
        return parts.sum(dim=0)
    out = torch.empty((N, K), dtype=torch.float32, device=parts.device)
    with torch.cuda.device(parts.device):
        _lib.call("nesie_gemm_sum_partials", ns, N * K, _lib.ptr(parts), _lib.ptr(out), _lib.stream())
    return out


def supported(n, k):
    return 1 <= n <= 1024 and k >= 1


class _LinearRows(Function):

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return gemm_nt(x, w)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            if supported(w.shape[1], w.shape[0]):
                gx = gemm_nt(gy, w, transpose_w=True)   # dX = dY W : B = W^T (K_in x K_out)
            else:
                gx = gy @ w
        if ctx.needs_input_grad[1]:
            if x.shape[1] <= 512 and gy.shape[0] >= 1:
                gw = wgrad(gy, x.contiguous())
            else:
                gw = gy.t() @ x
        return gx, gw


def linear_rows(x, w):
    """x (R, K) @ w (N, K)^T with fp32 parity on the tensor cores (library GEMM beyond N = 1024)."""
    if not supported(w.shape[0], w.shape[1]) or not x.is_cuda:
        return torch.nn.functional.linear(x, w)
    return _LinearRows.apply(x, w)
