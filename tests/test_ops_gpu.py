"""Parity tests proper: the CUDA path, called through the python API -> C ABI, against
(1) the CPU oracle, (2) the reference's own kernels compiled for sm_100a (oracle/_ref) on the
same GPU, and (3) the committed golden fixtures.  Index outputs must be bit-exact; fp32 feature
outputs of gather/group/interpolate are bit-exact as well (same contraction), gradients are
compared with a tolerance (atomic accumulation order)."""
import numpy as np
import pytest
import torch

import nesie_b200 as nb
from nesie_b200.synthetic import make_batch
from oracle import cpu, ref_cuda

pytestmark = pytest.mark.gpu


def dev(t):
    return t.cuda()


def scene_xyz(B, N, seed):
    return make_batch(B, max(N, 8), seed0=seed)[0][:, :N, :3].contiguous()


# ---------------------------------------------------------------- FPS
@pytest.mark.parametrize("B,N,M", [(2, 1500, 64), (1, 4096, 128), (2, 37, 37), (1, 1, 1),
                                   (1, 3, 3), (1, 31, 8), (1, 33, 33), (2, 512, 256),
                                   (2, 1024, 512), (2, 2048, 1024), (1, 8192, 300),
                                   (1, 8193, 300), (2, 10000, 256), (1, 33000, 128),
                                   (1, 70000, 64)])
def test_fps_matches_oracle(B, N, M):
    xyz = scene_xyz(B, N, 10 + N % 7)
    want = cpu.furthest_point_sample(xyz, M)
    got = nb.furthest_point_sample(dev(xyz), M)
    assert got.dtype == torch.int32 and got.shape == (B, M)
    assert torch.equal(got.cpu(), want)


def test_fps_ties_on_integer_grid():
    rng = np.random.default_rng(5)
    for N in [64, 100, 1500, 2049, 5000, 9000, 20000]:
        xyz = torch.from_numpy(rng.integers(0, 5, (2, N, 3)).astype(np.float32))
        M = min(N, 150)
        assert torch.equal(nb.furthest_point_sample(dev(xyz), M).cpu(),
                           cpu.furthest_point_sample(xyz, M)), N


def test_fps_full_size_vs_reference_kernel_and_oracle():
    """BASELINE shape: 40000 -> 2048, batch 8 (reference kernel on the same GPU), and batch 2
    against the CPU oracle."""
    xyz = make_batch(8, 40000, seed0=0)[0][..., :3].contiguous()
    got = nb.furthest_point_sample(dev(xyz), 2048)
    if ref_cuda.available():
        assert torch.equal(got, ref_cuda.furthest_point_sample(dev(xyz), 2048))
    assert torch.equal(got[:2].cpu(), cpu.furthest_point_sample(xyz[:2], 2048))
    # size-independent properties: unique indices (no exact duplicates picked twice unless the
    # cloud is exhausted), first index 0, prefix property (m' < m gives a prefix)
    assert (got[:, 0] == 0).all()
    assert torch.equal(nb.furthest_point_sample(dev(xyz), 300), got[:, :300])


def test_fps_stress_shape_100k():
    xyz = make_batch(2, 100000, seed0=50)[0][..., :3].contiguous()
    got = nb.furthest_point_sample(dev(xyz), 512)
    if ref_cuda.available():
        assert torch.equal(got, ref_cuda.furthest_point_sample(dev(xyz), 512))
    assert torch.equal(got[:1].cpu(), cpu.furthest_point_sample(xyz[:1], 512))


def test_stress_shape_op_chain_vs_reference_kernels():
    """BASELINE configs[4] (SAQE stress): 100 000-point scenes, 4096 SA1 centres, 64 samples.  The whole
    SA1 operator chain -- FPS, centre gather, ball query, grouping -- bit-exact against the reference's
    own kernels on the same GPU, plus size-independent properties."""
    pts = make_batch(2, 100000, seed0=70)[0]
    xyz = dev(pts[..., :3].contiguous())
    feats = dev(pts[..., 3:].transpose(1, 2).contiguous())
    idx = nb.furthest_point_sample(xyz, 4096)
    assert (idx[:, 0] == 0).all()
    assert all(idx[b].unique().numel() >= 4096 - 40 for b in range(2))     # duplicates only at exact ties
    centres = nb.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    bq = nb.ball_query(0.0, 0.2, 64, xyz, centres)
    d2 = ((torch.gather(xyz, 1, bq.long().reshape(2, -1, 1).expand(-1, -1, 3)).view(2, 4096, 64, 3)
           - centres.unsqueeze(2)) ** 2).sum(-1)
    assert float(d2.max()) < 0.2 ** 2 * (1 + 1e-6)                          # every hit inside the ball
    assert (bq[..., 1:] >= bq[..., :1]).all()                              # slots start from the first hit
    g = nb.grouping_operation(feats, bq)
    assert torch.equal(g, torch.gather(feats, 2, bq.long().reshape(2, 1, -1)).view(2, 1, 4096, 64))
    if ref_cuda.available():
        assert torch.equal(idx, ref_cuda.furthest_point_sample(xyz, 4096))
        assert torch.equal(bq, ref_cuda.ball_query(0.0, 0.2, 64, xyz, centres))
        assert torch.equal(g, ref_cuda.grouping_operation(feats, bq))


def test_fps_generic_fallback_large_n():
    xyz = torch.rand(1, 140000, 3)
    got = nb.furthest_point_sample(dev(xyz), 40)
    assert torch.equal(got.cpu(), cpu.furthest_point_sample(xyz, 40))


def test_fps_with_dist():
    pts = torch.randn(2, 300, 6)
    d = ((pts[:, :, None] - pts[:, None]) ** 2).sum(-1).contiguous()
    assert torch.equal(nb.furthest_point_sample_with_dist(dev(d), 50).cpu(),
                       cpu.furthest_point_sample_with_dist(d, 50))


def test_points_sampler_ranges():
    xyz = scene_xyz(2, 2000, 3)
    sampler = nb.Points_Sampler([64, 32], ['D-FPS', 'D-FPS'], [500, -1])
    got = sampler(dev(xyz), None).cpu()
    a = cpu.furthest_point_sample(xyz[:, :500].contiguous(), 64)
    b = cpu.furthest_point_sample(xyz[:, 500:].contiguous(), 32) + 500
    assert torch.equal(got, torch.cat([a, b], 1))


# ---------------------------------------------------------------- ball query
@pytest.mark.parametrize("B,N,M,K,r0,r1", [(2, 1500, 64, 16, 0.0, 0.3), (1, 4096, 128, 32, 0.0, 0.4),
                                           (1, 1000, 100, 8, 0.1, 0.5), (2, 37, 37, 4, 0.0, 0.8),
                                           (1, 1, 1, 3, 0.0, 0.2), (1, 2500, 700, 64, 0.0, 0.2),
                                           (3, 2048, 1024, 32, 0.0, 0.4), (2, 1024, 512, 16, 0.0, 0.8),
                                           (1, 3000, 5, 200, 0.0, 5.0)])
def test_ball_query_matches_oracle(B, N, M, K, r0, r1):
    xyz = scene_xyz(B, N, 20 + K)
    centres = xyz[:, torch.randperm(N, generator=torch.Generator().manual_seed(N))[:M]].contiguous()
    want = cpu.ball_query(r0, r1, K, xyz, centres)
    got = nb.ball_query(r0, r1, K, dev(xyz), dev(centres))
    assert got.dtype == torch.int32
    assert torch.equal(got.cpu(), want)


def test_ball_query_empty_balls_and_far_centres():
    xyz = scene_xyz(1, 3000, 4)
    centres = torch.cat([xyz[:, :10], xyz[:, :10] + 100.0], 1).contiguous()
    got = nb.ball_query(0.0, 0.2, 16, dev(xyz), dev(centres)).cpu()
    assert torch.equal(got, cpu.ball_query(0.0, 0.2, 16, xyz, centres))
    assert (got[:, 10:] == 0).all()


def test_ball_query_full_size_vs_reference_kernel():
    xyz = make_batch(8, 40000, seed0=0)[0][..., :3].contiguous()
    idx = nb.furthest_point_sample(dev(xyz), 2048)
    centres = torch.gather(dev(xyz), 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    got = nb.ball_query(0.0, 0.2, 64, dev(xyz), centres)
    if ref_cuda.available():
        assert torch.equal(got, ref_cuda.ball_query(0.0, 0.2, 64, dev(xyz), centres))
    assert torch.equal(got[:1].cpu(), cpu.ball_query(0.0, 0.2, 64, xyz[:1], centres[:1].cpu()))
    # property: every returned index is inside the ball (or the row's first hit repeated)
    p = torch.gather(dev(xyz), 1, got.reshape(8, -1, 1).long().expand(-1, -1, 3)).reshape(8, 2048, 64, 3)
    d2 = ((p - centres[:, :, None]) ** 2).sum(-1)
    assert (d2 < 0.2 * 0.2 * 1.0001).all()
    assert (got[:, :, 1:] >= got[:, :, :1]).all()


# ---------------------------------------------------------------- gather / group
@pytest.mark.parametrize("B,C,N,M", [(2, 3, 1500, 64), (1, 1, 10, 10), (2, 131, 2048, 1024), (1, 7, 33, 5)])
def test_gather_points(B, C, N, M):
    f = torch.randn(B, C, N)
    idx = torch.randint(0, N, (B, M), dtype=torch.int32)
    fg = dev(f).requires_grad_(True)
    out = nb.gather_points(fg, dev(idx))
    assert torch.equal(out.detach().cpu(), cpu.gather_points(f, idx))
    g = torch.randn(B, C, M)
    out.backward(dev(g))
    assert torch.allclose(fg.grad.cpu(), cpu.gather_points_grad(g, idx, N), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C,N,M,K", [(2, 4, 1500, 64, 16), (1, 131, 2048, 256, 32), (1, 3, 50, 7, 3),
                                       (2, 1, 40000, 128, 64), (1, 259, 512, 64, 16)])
def test_grouping_operation(B, C, N, M, K):
    f = torch.randn(B, C, N)
    idx = torch.randint(0, N, (B, M, K), dtype=torch.int32)
    fg = dev(f).requires_grad_(True)
    out = nb.grouping_operation(fg, dev(idx))
    assert torch.equal(out.detach().cpu(), cpu.grouping_operation(f, idx))
    g = torch.randn(B, C, M, K)
    out.backward(dev(g))
    assert torch.allclose(fg.grad.cpu(), cpu.grouping_operation_grad(g, idx, N), rtol=1e-4, atol=1e-4)


def test_query_and_group_matches_unfused_reference_graph():
    from oracle import modules as om
    xyz = scene_xyz(2, 3000, 9)
    feats = torch.randn(2, 5, 3000)
    centres = xyz[:, :200].contiguous()
    for normalize in (True, False):
        grouper = nb.QueryAndGroup(0.3, 16, use_xyz=True, normalize_xyz=normalize)
        xg = dev(xyz).requires_grad_(True)
        cg = dev(centres).requires_grad_(True)
        fg = dev(feats).requires_grad_(True)
        got = grouper(xg, cg, fg)
        xc = xyz.clone().requires_grad_(True)
        cc = centres.clone().requires_grad_(True)
        fc = feats.clone().requires_grad_(True)
        want, _ = om.query_and_group(xc, cc, fc, 0.3, 16, normalize_xyz=normalize)
        assert torch.equal(got.detach().cpu(), want.detach())  # fp32 bit-exact
        g = torch.randn_like(want)
        want.backward(g)
        got.backward(dev(g))
        assert torch.allclose(fg.grad.cpu(), fc.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(xg.grad.cpu(), xc.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(cg.grad.cpu(), cc.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("pad_to", [1, 4])
def test_query_and_group_rows_layout_matches_channel_major(pad_to):
    """The row-major (GEMM) grouped tensor holds the same bits as the reference-layout one; padded
    columns are zero and receive no gradient."""
    xyz = scene_xyz(2, 3000, 11)
    feats = torch.randn(2, 6, 3000)
    centres = xyz[:, :128].contiguous()
    grouper = nb.QueryAndGroup(0.3, 16, use_xyz=True, normalize_xyz=True)
    xa, ca, fa = (dev(t).requires_grad_(True) for t in (xyz, centres, feats))
    xb, cb, fb = (dev(t).requires_grad_(True) for t in (xyz, centres, feats))
    want = grouper(xa, ca, fa)                                   # (B, 9, 128, 16)
    rows = grouper.forward_rows(xb, cb, fb, pad_to=pad_to)       # (B*128*16, ld)
    ld = -(-9 // pad_to) * pad_to
    assert rows.shape == (2 * 128 * 16, ld)
    assert torch.equal(rows[:, :9].view(2, 128, 16, 9).permute(0, 3, 1, 2), want)
    assert (rows[:, 9:] == 0).all()
    g = torch.randn_like(want)
    want.backward(g)
    grows = torch.full_like(rows, 7.0)                           # garbage in the padding columns
    grows[:, :9] = g.permute(0, 2, 3, 1).reshape(-1, 9)
    rows.backward(grows)
    for a, b in ((xa, xb), (ca, cb), (fa, fb)):
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-5)


def test_query_and_group_rows_four_float_rows():
    """SA1's rows (coordinates + one feature: 16 bytes) take the thread-per-row kernel: same bits."""
    xyz = scene_xyz(2, 3000, 12)
    feats = torch.randn(2, 1, 3000)
    centres = xyz[:, :256].contiguous()
    grouper = nb.QueryAndGroup(0.2, 64, use_xyz=True, normalize_xyz=True)
    want = grouper(dev(xyz), dev(centres), dev(feats))           # (B, 4, 256, 64)
    rows = grouper.forward_rows(dev(xyz), dev(centres), dev(feats), pad_to=4)
    assert rows.shape == (2 * 256 * 64, 4)
    assert torch.equal(rows.view(2, 256, 64, 4).permute(0, 3, 1, 2), want)


# ---------------------------------------------------------------- three_nn / interpolate
@pytest.mark.parametrize("B,n,m", [(2, 512, 256), (2, 1024, 512), (1, 7, 2), (1, 5, 1), (1, 3000, 2500),
                                   (8, 20001, 1000)])   # last: the 4-targets-per-thread path
def test_three_nn(B, n, m):
    t = scene_xyz(B, n, 30)
    s = scene_xyz(B, m, 31)
    s[:, -1] = s[:, 0]  # duplicate source: earliest index must win
    dist, idx = nb.three_nn(dev(t), dev(s))
    wd, wi = cpu.three_nn(t, s)
    assert torch.equal(idx.cpu(), wi)
    # d^2 is bit-exact; the python-side torch.sqrt of the reference runs on the GPU, whose sqrt
    # may differ from the CPU's in the last ulp
    assert torch.allclose(dist.cpu(), wd, rtol=3e-7, atol=0)
    if ref_cuda.available():
        rd, ri = ref_cuda.three_nn(dev(t), dev(s))
        assert torch.equal(dist, rd) and torch.equal(idx, ri)


@pytest.mark.parametrize("B,C,m,n", [(2, 256, 256, 512), (1, 3, 10, 33), (2, 256, 512, 1024)])
def test_three_interpolate(B, C, m, n):
    f = torch.randn(B, C, m)
    idx = torch.randint(0, m, (B, n, 3), dtype=torch.int32)
    w = torch.rand(B, n, 3)
    w = w / w.sum(-1, keepdim=True)
    fg = dev(f).requires_grad_(True)
    out = nb.three_interpolate(fg, dev(idx), dev(w))
    assert torch.equal(out.detach().cpu(), cpu.three_interpolate(f, idx, w))
    g = torch.randn(B, C, n)
    out.backward(dev(g))
    assert torch.allclose(fg.grad.cpu(), cpu.three_interpolate_grad(g, idx, w, m), rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------- golden fixtures + ref kernels
def test_against_golden_fixture(golden_ref):
    g = golden_ref
    for i in range(int(g["n_cases"])):
        xyz = torch.from_numpy(g[f"c{i}_xyz"]).cuda()
        m, k = int(g[f"c{i}_m"]), int(g[f"c{i}_nsample"])
        idx = nb.furthest_point_sample(xyz, m)
        assert np.array_equal(idx.cpu().numpy(), g[f"c{i}_fps"]), i
        centres = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        bq = nb.ball_query(float(g[f"c{i}_min_r"]), float(g[f"c{i}_max_r"]), k, xyz, centres)
        assert np.array_equal(bq.cpu().numpy(), g[f"c{i}_bq"]), i
        feats = torch.from_numpy(g[f"c{i}_feats"]).cuda()
        assert np.array_equal(nb.grouping_operation(feats, bq).cpu().numpy(), g[f"c{i}_grouped"])
        gathered = nb.gather_points(feats, idx)
        assert np.array_equal(gathered.cpu().numpy(), g[f"c{i}_gathered"])
        dist, i3 = nb.three_nn(xyz, centres)
        assert np.array_equal(i3.cpu().numpy(), g[f"c{i}_nn_idx"])
        assert np.array_equal(dist.cpu().numpy(), torch.sqrt(torch.from_numpy(g[f"c{i}_nn_dist2"]).cuda()).cpu().numpy())
        w = torch.from_numpy(g[f"c{i}_weight"]).cuda()
        assert np.array_equal(nb.three_interpolate(gathered, i3, w).cpu().numpy(), g[f"c{i}_interp"])


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
def test_c_oracle_pinned_against_reference_kernels_live():
    """The oracle itself vs the reference's compiled kernels on fresh random inputs."""
    for seed, (B, N, M, K) in enumerate([(2, 3000, 200, 16), (1, 5000, 512, 32), (2, 777, 100, 8)]):
        xyz = scene_xyz(B, N, 40 + seed)
        xg = xyz.cuda()
        idx = ref_cuda.furthest_point_sample(xg, M)
        assert torch.equal(idx.cpu(), cpu.furthest_point_sample(xyz, M))
        centres = torch.gather(xg, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        assert torch.equal(ref_cuda.ball_query(0.0, 0.25, K, xg, centres).cpu(),
                           cpu.ball_query(0.0, 0.25, K, xyz, centres.cpu()))
        d, i3 = ref_cuda.three_nn_dist2(xg, centres)
        wd, wi = cpu.three_nn_dist2(xyz, centres.cpu())
        assert torch.equal(i3.cpu(), wi) and torch.equal(d.cpu(), wd)
        f = torch.randn(B, 6, M)
        w = torch.rand(B, N, 3)
        assert torch.equal(ref_cuda.three_interpolate(f.cuda(), i3, w.cuda()).cpu(),
                           cpu.three_interpolate(f, wi, w))


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        nb.furthest_point_sample(torch.rand(1, 10, 3), 2)
    with pytest.raises(RuntimeError):
        nb.ball_query(0.0, 1.0, 4, torch.rand(1, 10, 3), torch.rand(1, 2, 3))


# ---------------------------------------------------------------- grid ball query
@pytest.mark.parametrize("B,N,M,K,r0,r1", [(2, 1500, 64, 16, 0.0, 0.3), (1, 4096, 128, 32, 0.0, 0.4),
                                           (1, 1000, 100, 8, 0.1, 0.5), (2, 37, 37, 4, 0.0, 0.8),
                                           (1, 1, 1, 3, 0.0, 0.2), (2, 20000, 700, 64, 0.0, 0.2),
                                           (1, 3000, 5, 200, 0.0, 5.0), (1, 5000, 300, 16, 0.0, 0.01)])
def test_ball_query_grid_matches_oracle(B, N, M, K, r0, r1, monkeypatch):
    monkeypatch.setenv("NESIE_BALL_QUERY", "grid")
    xyz = scene_xyz(B, N, 20 + K)
    centres = xyz[:, torch.randperm(N, generator=torch.Generator().manual_seed(N))[:M]].contiguous()
    centres[:, -1] += 50.0  # a centre far outside the cloud
    want = cpu.ball_query(r0, r1, K, xyz, centres)
    got = nb.ball_query(r0, r1, K, dev(xyz), dev(centres))
    assert torch.equal(got.cpu(), want)


def test_ball_query_grid_duplicates_overflow_and_full_size(monkeypatch):
    # every point identical: each centre has N hits -> ordered-scan fallback
    xyz = torch.ones(1, 3000, 3)
    centres = torch.ones(1, 40, 3)
    monkeypatch.setenv("NESIE_BALL_QUERY", "grid")
    got = nb.ball_query(0.0, 0.2, 32, dev(xyz), dev(centres)).cpu()
    assert torch.equal(got, cpu.ball_query(0.0, 0.2, 32, xyz, centres))
    # integer grid (heavy duplicates, many > HMAX hit lists)
    rng = np.random.default_rng(8)
    xyz = torch.from_numpy(rng.integers(0, 6, (2, 9000, 3)).astype(np.float32))
    centres = xyz[:, :200].contiguous()
    got = nb.ball_query(0.0, 1.5, 64, dev(xyz), dev(centres)).cpu()
    assert torch.equal(got, cpu.ball_query(0.0, 1.5, 64, xyz, centres))
    # BASELINE shape: grid == brute force kernel == reference kernel
    xyz = make_batch(8, 40000, seed0=0)[0][..., :3].contiguous().cuda()
    idx = nb.furthest_point_sample(xyz, 2048)
    centres = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    grid = nb.ball_query(0.0, 0.2, 64, xyz, centres)
    monkeypatch.setenv("NESIE_BALL_QUERY", "brute")
    brute = nb.ball_query(0.0, 0.2, 64, xyz, centres)
    assert torch.equal(grid, brute)
    if ref_cuda.available():
        assert torch.equal(grid, ref_cuda.ball_query(0.0, 0.2, 64, xyz, centres))


@pytest.mark.parametrize("case", ["room", "uniform", "clustered", "flat", "duplicates", "few", "outside"])
def test_three_nn_grid_is_bit_identical_to_brute_force(case):
    """Exact 3-NN through the uniform grid: same distances and indices as nesie_three_nn, including the
    earliest-index tie rule, targets outside the sources' bounding box and degenerate source sets."""
    from nesie_b200.interpolate import three_nn_grid
    g = torch.Generator().manual_seed(hash(case) % 1000)
    B, n, m = 3, 5000, 1024
    if case == "room":
        src, tgt = scene_xyz(B, m, 40), scene_xyz(B, n, 41)
    elif case == "uniform":
        src, tgt = torch.rand(B, m, 3, generator=g) * 6, torch.rand(B, n, 3, generator=g) * 6
    elif case == "clustered":
        src = torch.randn(B, m, 3, generator=g) * 0.05 + torch.randint(0, 3, (B, m, 3), generator=g).float() * 2
        tgt = torch.rand(B, n, 3, generator=g) * 5
    elif case == "flat":
        src = torch.rand(B, m, 3, generator=g) * 4
        src[..., 2] = 1.0
        tgt = torch.rand(B, n, 3, generator=g) * 4
    elif case == "duplicates":
        src = torch.rand(B, 8, 3, generator=g).repeat(1, m // 8, 1)        # every source 128 times
        tgt = torch.rand(B, n, 3, generator=g)
    elif case == "few":
        m = 2
        src, tgt = torch.rand(B, m, 3, generator=g), torch.rand(B, n, 3, generator=g)
    else:
        src = torch.rand(B, m, 3, generator=g)
        tgt = torch.rand(B, n, 3, generator=g) * 20 - 10
    src, tgt = dev(src.contiguous()), dev(tgt.contiguous())
    d0, i0 = nb.three_nn(tgt, src)
    d1, i1, ws = three_nn_grid(tgt, src)
    assert torch.equal(d0, d1)
    assert torch.equal(i0, i1)
    d2, i2, _ = three_nn_grid(tgt[:, :777].contiguous(), src, ws)          # reuse the binned sources
    assert torch.equal(d2, d0[:, :777]) and torch.equal(i2, i0[:, :777])
