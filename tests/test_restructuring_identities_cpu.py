"""The algebra behind the restructured MiniPointNet / SA layers (pool_rows.py,
gather_linear.py), checked in float64 on the CPU with plain torch against the reference formulation
(models/dense_heads/side_pooling_module.py:343-370, ops/pointnet_modules/point_sa_module.py:136-158,
191-211): the kernels compute the right-hand sides, the reference the left-hand sides."""
import torch
import torch.nn.functional as F

torch.manual_seed(0)
D = torch.float64


def test_conv_over_concat_of_group_max_is_a_per_group_bias():
    """cat([max.expand, f]) @ W^T == f @ W_f^T + (max @ W_g^T)[group]  (+ its backward pieces)."""
    G, k, C, N = 7, 16, 12, 20
    y = torch.randn(G * k, C, dtype=D, requires_grad=True)
    b = torch.randn(C, dtype=D, requires_grad=True)
    w = torch.randn(N, 2 * C, dtype=D, requires_grad=True)
    f = (y + b).view(G, k, C)
    gmax, arg = f.max(dim=1)
    ref = torch.cat([gmax.unsqueeze(1).expand(-1, k, -1), f], dim=2).reshape(G * k, 2 * C) @ w.t()
    e = gmax @ w[:, :C].t() + w[:, C:] @ b                       # per-group vector (G, N)
    out = y @ w[:, C:].t() + e.repeat_interleave(k, dim=0)
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    g = torch.randn_like(ref)
    gy, gb, gw = torch.autograd.grad(ref, (y, b, w), g)
    # what _ConcatGlobalLinear.backward computes
    d_e = g.view(G, k, N).sum(1)
    d_y = g @ w[:, C:]
    d_g = d_e @ w[:, :C]
    d_y = d_y.view(G, k, C).scatter_add(1, arg.unsqueeze(1), d_g.unsqueeze(1)).reshape(G * k, C)
    colsum = d_e.sum(0)
    d_w = torch.cat([d_e.t() @ gmax, g.t() @ y + torch.outer(colsum, b)], dim=1)
    d_b = d_g.sum(0) + w[:, C:].t() @ colsum
    for a, r in ((d_y, gy), (d_b, gb), (d_w, gw)):
        assert torch.allclose(a.detach(), r, rtol=1e-10, atol=1e-10)


def test_bias_in_front_of_a_batchnorm_has_no_gradient():
    """zero_mean_grad: a BatchNorm's backward returns gradients with zero column sums, so the bias
    terms of the layer in front of it vanish."""
    x = torch.randn(64, 5, dtype=D, requires_grad=True)
    g = torch.randn(64, 5, dtype=D)
    out = F.batch_norm(x, None, None, torch.randn(5, dtype=D), torch.randn(5, dtype=D), True, 0.1, 1e-5)
    (gx,) = torch.autograd.grad(out, x, g)
    assert gx.sum(0).abs().max() < 1e-12


def test_first_conv_commutes_with_the_interpolation():
    """rows = [head | sum_j w_j f[idx_j]]: rows @ W^T == sum_j w_j (f @ W_f^T)[idx_j] + head @ W_x^T."""
    M, n, C, N = 30, 50, 9, 8
    feats = torch.randn(M, C, dtype=D)
    idx = torch.randint(0, M, (n, 3))
    wt = torch.rand(n, 3, dtype=D)
    head = torch.randn(n, 3, dtype=D)
    w = torch.randn(N, 3 + C, dtype=D)
    rows = torch.cat([head, (feats[idx] * wt.unsqueeze(-1)).sum(1)], dim=1)
    table = feats @ w[:, 3:].t()
    out = (table[idx] * wt.unsqueeze(-1)).sum(1) + head @ w[:, :3].t()
    assert torch.allclose(out, rows @ w.t(), rtol=1e-12, atol=1e-12)


def test_first_sa_conv_commutes_with_the_grouping():
    """rows = [(xyz[idx] - centre) / r | f[idx]]: rows @ W1^T == (f @ W1[:, 3:]^T)[idx] + rel @ W1[:, :3]^T,
    and the gradient w.r.t. the source features is the scatter of d_y W1[:, 3:] over idx."""
    Npts, M, K, C, N = 40, 6, 4, 5, 7
    xyz = torch.randn(Npts, 3, dtype=D)
    f = torch.randn(Npts, C, dtype=D, requires_grad=True)
    ctr = xyz[:M]
    idx = torch.randint(0, Npts, (M, K))
    w1 = torch.randn(N, 3 + C, dtype=D, requires_grad=True)
    rel = (xyz[idx] - ctr.unsqueeze(1)) / 0.4
    rows = torch.cat([rel, f[idx]], dim=2).reshape(M * K, 3 + C)
    ref = rows @ w1.t()
    table = f @ w1[:, 3:].t()
    out = table[idx.reshape(-1)] + rel.reshape(-1, 3) @ w1[:, :3].t()
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
    g = torch.randn_like(ref)
    gf, gw = torch.autograd.grad(ref, (f, w1), g)
    d_table = torch.zeros(Npts, N, dtype=D).index_add(0, idx.reshape(-1), g)       # gather_linear backward
    assert torch.allclose(d_table @ w1[:, 3:].detach(), gf, rtol=1e-10, atol=1e-10)
    d_w = torch.cat([g.t() @ rel.reshape(-1, 3), d_table.t() @ f.detach()], dim=1)
    assert torch.allclose(d_w, gw, rtol=1e-10, atol=1e-10)


def test_pooled_batchnorm_relu_from_group_extrema():
    """max over a group of relu(scale * y + shift) == relu(scale * (max y if scale >= 0 else min y) + shift)."""
    G, k, C = 9, 16, 6
    y = torch.randn(G, k, C, dtype=D)
    scale = torch.tensor([1.5, -0.7, 0.0, 2.0, -3.0, 0.2], dtype=D)
    shift = torch.randn(C, dtype=D)
    ref = torch.relu(y * scale + shift).max(dim=1).values
    ext = torch.where(scale >= 0, y.max(dim=1).values, y.min(dim=1).values)
    assert torch.equal(torch.relu(ext * scale + shift), ref)


def test_gradients_of_a_max_pooled_convolution_are_row_gathers():
    """out[g, c] = max_j (a[g k + j] . w[c]) + b[c]:  d_w[c] = sum_g d[g, c] a[g k + arg[g, c]]  (pool_wgrad),
    d_a[g k + j] = sum_{c: arg[g, c] = j} d[g, c] w[c]  (pool_dgrad / scatter + GEMM), d_b = column sums."""
    G, k, K, N = 8, 16, 10, 12
    a = torch.randn(G * k, K, dtype=D, requires_grad=True)
    w = torch.randn(N, K, dtype=D, requires_grad=True)
    b = torch.randn(N, dtype=D, requires_grad=True)
    yv = (a @ w.t()).view(G, k, N)
    out, arg = yv.max(dim=1)
    out = out + b
    d = torch.randn_like(out)
    ga, gw, gb = torch.autograd.grad(out, (a, w, b), d)
    picked = a.detach().view(G, k, K).gather(1, arg.unsqueeze(-1).expand(-1, -1, K))        # (G, N, K)
    assert torch.allclose((d.unsqueeze(-1) * picked).sum(0), gw, rtol=1e-10, atol=1e-10)
    dy = torch.zeros(G, k, N, dtype=D).scatter(1, arg.unsqueeze(1), d.unsqueeze(1))
    assert torch.allclose(dy.view(G * k, N) @ w.detach(), ga, rtol=1e-10, atol=1e-10)
    assert torch.allclose(d.sum(0), gb, rtol=1e-12, atol=1e-12)


def test_unit_maxima_combine_to_group_maxima_with_the_first_row_rule():
    """Groups of 64 rows arrive as two 32-row units from the epilogue; strict '>' keeps the first row."""
    G, C = 11, 5
    y = torch.randint(-3, 4, (G, 64, C)).double()                       # many exact ties
    u_max, u_arg = y.view(G, 2, 32, C).max(dim=2)
    # first maximising row inside each unit (torch.max's index on ties is unspecified: recompute)
    u_arg = (y.view(G, 2, 32, C) == u_max.unsqueeze(2)).double().argmax(dim=2)
    best, arg = u_max[:, 0].clone(), u_arg[:, 0].clone()
    take = u_max[:, 1] > best
    best[take], arg[take] = u_max[:, 1][take], (32 + u_arg[:, 1])[take]
    ref = y.max(dim=1).values
    first = (y == ref.unsqueeze(1)).double().argmax(dim=1)
    assert torch.equal(best, ref) and torch.equal(arg, first)
