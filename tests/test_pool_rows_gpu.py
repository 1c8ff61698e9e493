"""Row GEMMs with the group maximum / a per-group bias in the epilogue (nesie_gemm_nt_3xtf32_pool) and the
MiniPointNet built on them (pool_rows.py) against plain torch fp32 formulations of
models/dense_heads/side_pooling_module.py:343-370 and against the step-by-step kernels."""
import pytest
import torch

from nesie_b200 import _lib
from nesie_b200 import pool_rows
from nesie_b200.linear_rows import _pack
from nesie_b200.side_pooling import MiniPointNet, SidePooling

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,K,N,k", [(4096, 256, 128, 16), (4096 + 64, 128, 256, 32), (2048, 64, 128, 64),
                                     (1600, 260, 132, 16), (65536 + 128, 256, 128, 16),   # CTA-pair launch,
                                     (40000 - 32, 128, 256, 32)])                         # odd tile count
def test_gemm_pool_epilogue(R, K, N, k):
    """Unit maxima + first maximising row, with and without storing the output; integers so that the
    3xTF32 product is exact and ties are real ties."""
    torch.manual_seed(R + k)
    x = torch.randint(-3, 4, (R, K), device="cuda").float()
    w = torch.randint(-2, 3, (N, K), device="cuda").float()
    img = _pack(w, N, K, K, 1)
    want = x @ w.t()
    for store in (True, False):
        y, _, pmax, amax = pool_rows._gemm_pool(x, img, N, None, None, False, store, k)
        if store:
            assert torch.equal(y, want)
        else:
            assert y is None
        out, arg = pool_rows._finalize(pmax, amax, None, k)
        wmax, warg = want.view(R // k, k, N).max(dim=1)
        assert torch.equal(out, wmax)
        # first maximising row (torch.max's index on CUDA is not specified for ties: compare by value
        # and check minimality separately)
        rows = want.view(R // k, k, N)
        assert torch.equal(rows.gather(1, arg.long().unsqueeze(1)).squeeze(1), wmax)
        first = (rows == wmax.unsqueeze(1)).float().argmax(dim=1)
        assert torch.equal(arg.long(), first)


def test_gemm_group_bias_and_stats():
    torch.manual_seed(3)
    R, K, N, k = 4096, 128, 256, 16
    x = torch.randn(R, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.1
    e = torch.randn(R // k, N, device="cuda")
    img = _pack(w, N, K, K, 1)
    y, parts, _, _ = pool_rows._gemm_pool(x, img, N, None, None, True, True, 0, grp_bias=e, grp_k=k)
    want = (x.double() @ w.double().t() + e.double().repeat_interleave(k, dim=0))
    assert (y.double() - want).abs().max() < 2e-5
    sums = parts.double().sum(dim=0)
    assert torch.allclose(sums[0], want.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(sums[1], (want * want).sum(0), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("G,boxes,cin", [(16, 96, 259), (64, 40, 131)])
def test_mini_pointnet_pooled_matches_tensor_formulation(G, boxes, cin, monkeypatch):
    """Outputs, input-free parameter gradients and running statistics of the pooled path against the
    module's own nn.Conv2d / BatchNorm2d / max / cat forward (the reference formulation) in fp32."""
    torch.manual_seed(G)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)     # the twin's convolutions in fp32
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    mpn = MiniPointNet(cin, 128).cuda()
    twin = MiniPointNet(cin, 128).cuda()
    twin.load_state_dict(mpn.state_dict())
    ld = -(-cin // 4) * 4
    rows = torch.randn(boxes * G, ld, device="cuda")
    rows[:, cin:] = 0
    got = SidePooling._mini_pointnet_pooled(mpn, rows, G)
    assert got is not None
    x4 = rows[:, :cin].view(1, boxes, G, cin).permute(0, 3, 1, 2).contiguous()     # (1, C, boxes, G)
    want = twin(x4)[0].t()                                                          # (boxes, 128)
    assert (got - want).abs().max() < 2e-5 * want.abs().max().clamp_min(1.0)
    g = torch.randn_like(want)
    (got * g).sum().backward()
    (want * g).sum().backward()
    gmax = max(float(p.grad.norm()) for p in twin.parameters())
    for (name, p), q in zip(twin.named_parameters(), mpn.parameters()):
        err = float((q.grad - p.grad).norm())
        assert err < 5e-3 * float(p.grad.norm()) or err < 1e-5 * gmax, (name, err, float(p.grad.norm()))
    for (name, b1), b2 in zip(twin.named_buffers(), mpn.buffers()):
        assert torch.allclose(b1.float(), b2.float(), rtol=1e-4, atol=1e-5), name


def test_pooled_path_equals_stepwise_kernels(monkeypatch):
    """Same weights through the pooled path and through the step-by-step kernels (NESIE_POOL_FUSE=0)."""
    torch.manual_seed(9)
    sp = SidePooling(18, 1, 18, None, 8, "vote", seed_feat_dim=256).cuda()
    mpn = sp.mlps_before[0]
    rows = torch.randn(128 * 16, 260, device="cuda")
    rows[:, 259:] = 0
    a = sp._mini_pointnet(mpn, rows, 16)
    ga = torch.autograd.grad((a * a).sum(), list(mpn.parameters()))
    monkeypatch.setenv("NESIE_POOL_FUSE", "0")
    b = sp._mini_pointnet(mpn, rows, 16)
    gb = torch.autograd.grad((b * b).sum(), list(mpn.parameters()))
    assert (a - b).abs().max() < 1e-5 * b.abs().max()
    top = max(float(g.norm()) for g in gb)
    for (name, _), x, y in zip(mpn.named_parameters(), ga, gb):
        err = float((x - y).norm())
        assert err < 2e-3 * float(y.norm()) or err < 1e-5 * top, (name, err, float(y.norm()))


def test_pool_wgrad_kernel_is_a_row_gather():
    torch.manual_seed(1)
    G, k, N, K = 300, 16, 128, 256
    y = torch.randn(G * k, K, device="cuda")
    sc, sh = torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda") * 0.3
    d = torch.randn(G, N, device="cuda")
    arg = torch.randint(0, k, (G, N), device="cuda", dtype=torch.uint8)
    parts = torch.empty((_lib.lib().nesie_pool_wgrad_parts(G), N, K), device="cuda")
    _lib.call("nesie_pool_wgrad", G, k, N, K, _lib.ptr(d), _lib.ptr(arg), _lib.ptr(y), _lib.ptr(sc),
              _lib.ptr(sh), _lib.ptr(parts), _lib.stream())
    a = torch.relu(y * sc + sh).view(G, k, K).double()
    picked = a.gather(1, arg.long().unsqueeze(-1).expand(-1, -1, K))              # (G, N, K)
    want = (d.double().unsqueeze(-1) * picked).sum(0)
    assert (parts.double().sum(0) - want).abs().max() < 1e-4


@pytest.mark.parametrize("rows,n", [(4096, 256), (4096, 128), (37, 4), (100000, 1024)])
def test_colsum_rows(rows, n):
    torch.manual_seed(rows)
    x = torch.randn(rows, n, device="cuda")
    got = pool_rows._colsum(x)
    assert torch.allclose(got.double(), x.double().sum(0), rtol=1e-5, atol=1e-3)
    assert torch.equal(got, pool_rows._colsum(x))        # fixed summation order


@pytest.mark.parametrize("k,N,K", [(16, 128, 256), (64, 128, 256), (5, 36, 520), (16, 256, 128)])
def test_pool_dgrad_kernel(k, N, K):
    torch.manual_seed(k)
    G = 200
    d = torch.randn(G, N, device="cuda")
    arg = torch.randint(0, k, (G, N), device="cuda", dtype=torch.uint8)
    w = torch.randn(N, K, device="cuda")
    got = torch.empty(G * k, K, device="cuda")
    _lib.call("nesie_pool_dgrad", G, k, N, K, _lib.ptr(d), _lib.ptr(arg), _lib.ptr(w), _lib.ptr(got),
              _lib.stream())
    dy = torch.zeros(G, k, N, device="cuda", dtype=torch.float64)
    dy.scatter_(1, arg.long().unsqueeze(1), d.double().unsqueeze(1))
    want = dy.view(G * k, N) @ w.double()
    assert (got.double() - want).abs().max() < 1e-4


@pytest.mark.parametrize("zero_mean", [False, True])
def test_concat_global_linear_matches_torch(zero_mean):
    """cat([max.expand, y + b]) @ W^T through the group-bias GEMM, forward and every gradient; with
    zero_mean_grad the output gradient is centred first (as a BatchNorm backward delivers it)."""
    torch.manual_seed(5)
    G, k, C, N = 256, 16, 128, 256
    y = torch.randn(G * k, C, device="cuda", requires_grad=True)
    b = torch.randn(C, device="cuda", requires_grad=True)
    w = (torch.randn(N, 2 * C, device="cuda") * 0.1).requires_grad_(True)
    f = (y + b).view(G, k, C)
    gmax, arg = f.max(dim=1)
    got, _ = pool_rows.concat_global_linear(y, gmax.detach(), arg.to(torch.uint8), b, w, k, zero_mean)
    yd, bd, wd = (t.detach().double().requires_grad_(True) for t in (y, b, w))
    fd = (yd + bd).view(G, k, C)
    want = torch.cat([fd.max(dim=1).values.unsqueeze(1).expand(-1, k, -1), fd], dim=2).reshape(G * k, 2 * C) @ wd.t()
    assert (got.double() - want).abs().max() < 2e-5
    g = torch.randn(G * k, N, device="cuda")
    if zero_mean:
        g = g - g.mean(dim=0, keepdim=True)
    got.backward(g)
    want.backward(g.double())
    for a, r in ((y, yd), (w, wd)):
        assert (a.grad.double() - r.grad).abs().max() < 1e-4 * r.grad.abs().max().clamp_min(1.0)
    if zero_mean:     # a constant in front of centred gradients: analytically zero, returned as zero
        assert float(b.grad.abs().max()) == 0.0 and float(bd.grad.abs().max()) < 1e-3
    else:
        assert (b.grad.double() - bd.grad).abs().max() < 1e-4 * bd.grad.abs().max().clamp_min(1.0)


@pytest.mark.parametrize("J,C,with_head", [(3, 256, True), (1, 128, False), (3, 64, True), (1, 128, True)])
def test_gather_linear_matches_torch(J, C, with_head):
    from nesie_b200.gather_linear import gather_linear
    torch.manual_seed(J * C)
    B, M, n = 3, 200, 1000
    table = torch.randn(B, M, C, device="cuda", requires_grad=True)
    idx = torch.randint(0, M, (B, n, J), device="cuda", dtype=torch.int32)
    w = torch.rand(B, n, J, device="cuda") if J == 3 else None
    head = torch.randn(B, n, 3, device="cuda") if with_head else None
    wx = torch.randn(C, 3, device="cuda", requires_grad=True) if with_head else None
    y, parts = gather_linear(table, idx, w, head, wx, True)
    td = table.detach().double().requires_grad_(True)
    rows = torch.stack([td[b][idx[b].long()] for b in range(B)])            # (B, n, J, C)
    want = (rows * (w.double().unsqueeze(-1) if w is not None else 1.0)).sum(2)
    if with_head:
        wxd = wx.detach().double().requires_grad_(True)
        want = want + head.double() @ wxd.t()
    want = want.reshape(B * n, C)
    assert (y.double() - want).abs().max() < 1e-5
    sums = parts.double().sum(0)
    assert torch.allclose(sums[0], want.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(sums[1], (want * want).sum(0), rtol=1e-5, atol=1e-3)
    g = torch.randn(B * n, C, device="cuda")
    y.backward(g)
    want.backward(g.double())
    assert (table.grad.double() - td.grad).abs().max() < 1e-4
    if with_head:
        assert (wx.grad.double() - wxd.grad).abs().max() < 1e-3


def test_factored_first_layer_equals_dense_rows(monkeypatch):
    """SidePooling forward + parameter gradients with the first convolutions commuted with the
    interpolation (gather_linear) against the dense-rows GEMM path."""
    import copy
    torch.manual_seed(2)
    B, K, N, C = 2, 16, 128, 256
    a = SidePooling(18, 1, 18, None, K // 2, "vote", seed_feat_dim=C).cuda()
    b = copy.deepcopy(a)
    g = torch.Generator().manual_seed(3)
    center = (torch.rand(B, K, 3, generator=g) * 4 - 2).cuda()
    size = (torch.rand(B, K, 3, generator=g) * 1.5 + 0.2).cuda()
    heading = ((torch.rand(B, K, generator=g) - 0.5)).cuda()
    ep = {"seed_points": (torch.rand(B, N, 3, generator=g) * 5 - 2.5).cuda(),
          "seed_features": torch.randn(B, C, N, generator=g).cuda(),
          "bbox_probs": torch.softmax(torch.randn(B, 6, 33, K // 2, generator=g), dim=2).cuda()}
    outs = []
    for mod, flag in ((a, "1"), (b, "0")):
        monkeypatch.setenv("NESIE_GATHER_LINEAR", flag)
        o = mod(center, size, heading, dict(ep))
        ((o["side_scores"] ** 2).sum() + (o["iou_scores"] ** 2).sum()).backward()
        outs.append(o)
    for key in ("side_scores", "iou_scores"):
        assert (outs[0][key] - outs[1][key]).abs().max() < 2e-5 * outs[1][key].abs().max().clamp_min(1.0), key
    top = max(float(p.grad.norm()) for p in b.parameters())
    for (name, p), q in zip(b.named_parameters(), a.parameters()):
        err = float((q.grad - p.grad).norm())
        assert err < 5e-3 * float(p.grad.norm()) or err < 1e-5 * top, (name, err, float(p.grad.norm()))


def test_gather_linear_sa_grouping_head():
    """xyz + center mode: head = (xyz[idx] - center) / radius, as nesie_group_rows builds the rows."""
    from nesie_b200.gather_linear import gather_linear
    torch.manual_seed(8)
    B, M, npoint, ns, C = 2, 300, 40, 16, 128
    table = torch.randn(B, M, C, device="cuda")
    xyz = torch.randn(B, M, 3, device="cuda")
    center = torch.randn(B, npoint, 3, device="cuda")
    idx = torch.randint(0, M, (B, npoint * ns, 1), device="cuda", dtype=torch.int32)
    wx = torch.randn(C, 3, device="cuda", requires_grad=True)
    y, _ = gather_linear(table, idx, None, None, wx, True, xyz, center, ns, 0.4)
    li = idx.long().squeeze(-1)
    head = (torch.stack([xyz[b][li[b]] for b in range(B)]) - center.repeat_interleave(ns, dim=1)) * (1.0 / 0.4)
    want = torch.stack([table[b][li[b]] for b in range(B)]).double() + head.double() @ wx.detach().double().t()
    assert (y.double() - want.reshape(-1, C)).abs().max() < 1e-5
    g = torch.randn_like(y)
    y.backward(g)
    want_wx = g.double().t() @ head.double().reshape(-1, 3)
    assert (wx.grad.double() - want_wx).abs().max() < 1e-3


def test_sa_module_commuted_first_layer_equals_grouped_rows(monkeypatch):
    from nesie_b200.pointnet_modules import PointSAModule
    torch.manual_seed(6)
    sa = PointSAModule(num_point=64, radius=0.4, num_sample=16, mlp_channels=[128, 128, 128, 256],
                       use_xyz=True, normalize_xyz=True).cuda()
    xyz = torch.rand(2, 512, 3, device="cuda")
    feats = torch.randn(2, 128, 512, device="cuda", requires_grad=True)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("NESIE_GATHER_LINEAR", flag)
        sa.zero_grad()
        feats.grad = None
        _, out, _ = sa(xyz, feats)
        (out * out).sum().backward()
        res.append((out.detach().clone(), feats.grad.clone(), [p.grad.clone() for p in sa.parameters()]))
    (o1, f1, p1), (o0, f0, p0) = res
    assert (o1 - o0).abs().max() < 2e-5 * o0.abs().max()
    assert (f1 - f0).norm() < 2e-3 * f0.norm()
    top = max(float(g.norm()) for g in p0)
    for (name, _), a, b in zip(sa.named_parameters(), p1, p0):
        err = float((a - b).norm())
        assert err < 5e-3 * float(b.norm()) or err < 1e-5 * top, (name, err, float(b.norm()))


@pytest.mark.parametrize("k,neg", [(16, False), (32, True), (64, False)])
def test_sa_pooled_last_layer_from_gemm_epilogue(k, neg, monkeypatch):
    """BatchNorm + ReLU + max-pool of an SA level from the GEMM epilogue's unit maxima / minima
    (negative BatchNorm weights exercise the minimum) against the kernel that re-reads the pre-activation."""
    from nesie_b200.pointnet_modules import PointSAModule
    torch.manual_seed(k)
    sa = PointSAModule(num_point=48, radius=0.5, num_sample=k, mlp_channels=[32, 64, 64, 128],
                       use_xyz=True, normalize_xyz=True).cuda()
    if neg:
        with torch.no_grad():
            for m in sa.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.weight.copy_(torch.randn_like(m.weight))
    xyz = torch.rand(2, 600, 3, device="cuda")
    feats = torch.randn(2, 32, 600, device="cuda", requires_grad=True)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("NESIE_POOL_FUSE", flag)
        sa.zero_grad()
        feats.grad = None
        _, out, _ = sa(xyz, feats)
        (out * out).sum().backward()
        res.append((out.detach().clone(), feats.grad.clone(), [p.grad.clone() for p in sa.parameters()]))
    (o1, f1, p1), (o0, f0, p0) = res
    assert torch.equal(o1, o0)                      # same GEMM, same statistics, max of a monotone map
    assert (f1 - f0).norm() < 1e-4 * f0.norm()
    for a, b in zip(p1, p0):
        assert (a - b).norm() <= 1e-4 * b.norm() + 1e-7
