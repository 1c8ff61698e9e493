"""Shared MLP with BatchNorm fused into the tcgen05 GEMMs (nesie_b200/mlp_rows.py) against a float64
torch restatement of conv -> BN(batch stats) -> ReLU (ops/pointnet_modules/point_sa_module.py:279-288)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from nesie_b200 import mlp_rows

pytestmark = pytest.mark.gpu


def _layers(chs, dtype, seed):
    torch.manual_seed(seed)
    out = []
    for k, n in zip(chs[:-1], chs[1:]):
        w = (torch.randn(n, k) * (2.0 / k) ** 0.5).to("cuda", dtype).requires_grad_(True)
        bn = nn.BatchNorm1d(n).to("cuda", dtype)
        with torch.no_grad():
            bn.weight.copy_(torch.rand(n) + 0.5)
            bn.bias.copy_(torch.randn(n) * 0.3)
        out.append((w, bn))
    return out


def _reference(x, layers, pool_k):
    for w, bn in layers:
        x = F.relu(F.batch_norm(F.linear(x, w), bn.running_mean, bn.running_var, bn.weight, bn.bias,
                                True, bn.momentum, bn.eps))
    if pool_k:
        x = x.view(-1, pool_k, x.shape[1]).amax(dim=1)
    return x


@pytest.mark.parametrize("R,chs,pool_k", [(4096, (8, 64, 64, 128), 16), (3000, (132, 128, 128, 256), 0),
                                          (129 * 32, (4, 64, 64, 128), 32), (640, (256, 128, 128), 0),
                                          (70000, (64, 64), 0)])
def test_fused_mlp_matches_float64(R, chs, pool_k):
    layers = _layers(chs, torch.float32, 1)
    ref = _layers(chs, torch.float64, 1)
    torch.manual_seed(5)
    x = (torch.randn(R, chs[0], device="cuda") + 0.5).requires_grad_(True)
    xd = x.detach().double().requires_grad_(True)
    assert mlp_rows.supported(x, layers)
    got = mlp_rows.mlp_rows(x, layers, pool_k)
    want = _reference(xd, ref, pool_k)
    scale = want.abs().max().item()
    assert (got.double() - want).abs().max().item() < 2e-5 * scale
    g = torch.randn_like(got)
    got.backward(g)
    want.backward(g.double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()  # noqa: E731
    # A pre-activation within fp32 rounding of zero takes the other branch of the ReLU derivative
    # than in float64: with millions of elements that happens to a row now and then, and it is
    # not an error of the kernels.  Such rows are counted, not compared.
    row_err = (x.grad.double() - xd.grad).abs().max(dim=1).values / xd.grad.abs().max()
    flipped = int((row_err > 1e-4).sum())
    assert flipped <= R // 20000, (flipped, row_err.max().item())
    ptol = 1e-4 if flipped == 0 else 1e-2
    for (w, bn), (wd, bnd) in zip(layers, ref):
        assert rel(w.grad, wd.grad) < ptol
        assert rel(bn.weight.grad, bnd.weight.grad) < ptol
        assert rel(bn.bias.grad, bnd.bias.grad) < ptol
        assert rel(bn.running_mean, bnd.running_mean) < 1e-5
        assert rel(bn.running_var, bnd.running_var) < 1e-5
        assert int(bn.num_batches_tracked) == 1


def test_bn_backward_statistics_in_the_dgrad_epilogue(monkeypatch):
    """NESIE_BNBWD_FUSE=1: the BatchNorm-backward column sums come from the data-gradient GEMM's
    epilogue (nesie_gemm_nt_3xtf32_bnbwd) instead of a sweep over (g, y): same gradients as the
    default path (only the order of the fp32 column sums differs)."""
    chs, R = (132, 128, 128, 256), 3000
    torch.manual_seed(7)
    x0 = torch.randn(R, chs[0], device="cuda")
    g = torch.randn(R, chs[-1], device="cuda")
    grads = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NESIE_BNBWD_FUSE", mode)
        layers = _layers(chs, torch.float32, 1)
        x = x0.clone().requires_grad_(True)
        mlp_rows.mlp_rows(x, layers).backward(g)
        grads[mode] = [x.grad] + [t.grad for w, bn in layers for t in (w, bn.weight, bn.bias)]
    for a, b in zip(grads["0"], grads["1"]):
        assert float((a - b).abs().max()) < 1e-4 * float(a.abs().max())


def test_fused_mlp_with_large_mean_keeps_variance():
    """The GEMM epilogue sums y and y^2 without a pivot: a mean ten times the spread must still give
    the variance to 1e-4."""
    layers = _layers((64, 64), torch.float32, 2)
    ref = _layers((64, 64), torch.float64, 2)
    x = torch.randn(20000, 64, device="cuda") * 0.1 + 3.0
    got = mlp_rows.mlp_rows(x, layers)
    want = _reference(x.double(), ref, 0)
    assert (got.double() - want).abs().max().item() < 5e-4 * want.abs().max().item()
    assert ((layers[0][1].running_var.double() - ref[0][1].running_var).abs().max() /
            ref[0][1].running_var.abs().max()).item() < 1e-4


def test_unsupported_shapes_are_reported():
    layers = _layers((7, 64), torch.float32, 3)
    assert not mlp_rows.supported(torch.randn(100, 7, device="cuda"), layers)


def test_fused_entry_points_refuse_what_they_cannot_do():
    """The fused C entry points fail loudly (status + message) instead of falling back."""
    import ctypes
    from nesie_b200 import _lib
    from nesie_b200.linear_rows import _pack
    R, K, N = 256, 6, 64                       # K not a multiple of 4: no TMA path for the operand
    a = torch.randn(R, K, device="cuda")
    w = torch.randn(N, K, device="cuda")
    out = torch.empty(R, N, device="cuda")
    img = _pack(w, N, K, K, 1)
    sc = torch.ones(K, device="cuda")
    assert not _lib.lib().nesie_gemm_fused_supported(R, N, K, ctypes.c_void_p(a.data_ptr()), K, N)
    with pytest.raises(RuntimeError, match="not supported"):
        _lib.call("nesie_gemm_nt_3xtf32_fused", R, N, K, _lib.ptr(a), K, _lib.ptr(img), _lib.ptr(out), N,
                  _lib.ptr(sc), _lib.ptr(sc), None, _lib.stream())
    gy = torch.randn(R, N, device="cuda")
    ns = _lib.lib().nesie_gemm_wgrad_splits(R, N, K)
    parts = torch.empty(ns, N, K, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of 4"):
        _lib.call("nesie_gemm_wgrad_3xtf32_fused", R, N, K, _lib.ptr(gy), N, _lib.ptr(a), K,
                  _lib.ptr(sc), _lib.ptr(sc), _lib.ptr(parts), ns, _lib.stream())
    # scale without shift
    a4 = torch.randn(R, 8, device="cuda")
    img4 = _pack(torch.randn(N, 8, device="cuda"), N, 8, 8, 1)
    with pytest.raises(RuntimeError, match="together"):
        _lib.call("nesie_gemm_nt_3xtf32_fused", R, N, 8, _lib.ptr(a4), 8, _lib.ptr(img4), _lib.ptr(out), N,
                  _lib.ptr(sc), None, None, _lib.stream())
