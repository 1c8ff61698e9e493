"""Module-level parity: PointSAModule / PointFPModule / PointNet2SASSG on the GPU (this repo's
kernels + cuDNN/cuBLAS fp32 MLPs, TF32 off) against the reference-shaped CPU forward of
oracle/modules.py with shared weights.  Indices bit-exact; features within 1e-5 of the feature
scale (fp32 accumulation order differs between cuDNN and the CPU conv); gradients 1e-4."""
import copy

import pytest
import torch

import nesie_b200 as nb
from nesie_b200.synthetic import make_batch
from oracle import modules as om

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def test_sa_module_forward_backward():
    torch.manual_seed(0)
    pts = make_batch(2, 4096, seed0=3)[0]
    xyz = pts[..., :3].contiguous()
    feats = torch.randn(2, 5, 4096)
    sa = nb.PointSAModule(mlp_channels=[5, 16, 16, 32], num_point=128, radius=0.4, num_sample=16,
                          use_xyz=True, normalize_xyz=True)
    sa_cpu = copy.deepcopy(sa)
    fc = feats.clone().requires_grad_(True)
    wx, wf, wi = om.sa_forward(sa_cpu, xyz, fc)
    sa = sa.cuda()
    fg = feats.cuda().requires_grad_(True)
    gx, gf, gi = sa(xyz.cuda(), fg)
    assert torch.equal(gi.cpu(), wi) and torch.equal(gx.cpu(), wx)
    assert rel_err(gf.detach().cpu(), wf.detach()) < 1e-5
    g = torch.randn_like(wf)
    wf.backward(g)
    gf.backward(g.cuda())
    assert rel_err(fg.grad.cpu(), fc.grad) < 1e-4
    for (n1, p1), (n2, p2) in zip(sa.named_parameters(), sa_cpu.named_parameters()):
        assert n1 == n2 and rel_err(p1.grad.cpu(), p2.grad) < 1e-3, n1


def test_sa_module_given_indices_and_target_xyz():
    pts = make_batch(1, 2048, seed0=5)[0]
    xyz = pts[..., :3].contiguous()
    sa = nb.PointSAModule(mlp_channels=[0, 8, 8], num_point=64, radius=0.5, num_sample=8)
    sa_cpu = copy.deepcopy(sa)
    sa = sa.cuda()
    idx = torch.randperm(2048)[:64].to(torch.int32)[None]
    gx, gf, gi = sa(xyz.cuda(), None, indices=idx.cuda())
    wx, wf, _ = om.sa_forward(sa_cpu, xyz, None, indices=idx)
    assert torch.equal(gx.cpu(), wx) and rel_err(gf.detach().cpu(), wf.detach()) < 1e-5
    tgt = xyz[:, :64].contiguous() + 0.01
    gx, gf, gi = sa(xyz.cuda(), None, target_xyz=tgt.cuda())
    wx, wf, _ = om.sa_forward(sa_cpu, xyz, None, target_xyz=tgt)
    assert gi is None and torch.equal(gx.cpu(), wx)
    assert rel_err(gf.detach().cpu(), wf.detach()) < 1e-5


def test_fp_module_forward_backward():
    torch.manual_seed(1)
    target, source = torch.rand(2, 512, 3), torch.rand(2, 256, 3)
    tf, sf = torch.randn(2, 24, 512), torch.randn(2, 32, 256)
    fp = nb.PointFPModule(mlp_channels=[56, 32, 32])
    fp_cpu = copy.deepcopy(fp)
    sc = sf.clone().requires_grad_(True)
    want = om.fp_forward(fp_cpu, target, source, tf, sc)
    fp = fp.cuda()
    sg = sf.cuda().requires_grad_(True)
    got = fp(target.cuda(), source.cuda(), tf.cuda(), sg)
    assert rel_err(got.detach().cpu(), want.detach()) < 1e-5
    g = torch.randn_like(want)
    want.backward(g)
    got.backward(g.cuda())
    assert rel_err(sg.grad.cpu(), sc.grad) < 1e-4


def test_fp_module_with_more_than_512_input_channels():
    """A 512 + 256 = 768-channel FP layer: wider than the weight-gradient kernel takes (k <= 512), so the
    fused GEMM + BatchNorm pipeline must decline it up front and forward AND backward must still work."""
    torch.manual_seed(5)
    target, source = torch.rand(2, 256, 3), torch.rand(2, 128, 3)
    tf, sf = torch.randn(2, 256, 256), torch.randn(2, 512, 128)
    fp = nb.PointFPModule(mlp_channels=[768, 256, 256])
    fp_cpu = copy.deepcopy(fp)
    sc = sf.clone().requires_grad_(True)
    want = om.fp_forward(fp_cpu, target, source, tf, sc)
    fp = fp.cuda()
    sg = sf.cuda().requires_grad_(True)
    got = fp(target.cuda(), source.cuda(), tf.cuda(), sg)
    assert rel_err(got.detach().cpu(), want.detach()) < 1e-5
    g = torch.randn_like(want)
    want.backward(g)
    got.backward(g.cuda())
    assert rel_err(sg.grad.cpu(), sc.grad) < 1e-4
    for (n1, p1), (n2, p2) in zip(fp.named_parameters(), fp_cpu.named_parameters()):
        assert n1 == n2 and rel_err(p1.grad.cpu(), p2.grad) < 1e-3, n1


@pytest.mark.parametrize("overlap", [True, False])
def test_backbone_votenet_shape(overlap):
    """PointNet2SASSG at a reduced ScanNet shape (8192 pts) end to end, train-mode BN."""
    torch.manual_seed(2)
    pts = make_batch(2, 8192, seed0=11)[0]
    bb = nb.PointNet2SASSG(in_channels=4, num_points=(512, 256, 128, 64), radius=(0.2, 0.4, 0.8, 1.2),
                           num_samples=(32, 16, 16, 16), overlap_fps=overlap)
    bb_cpu = copy.deepcopy(bb)
    want = om.backbone_forward(bb_cpu, pts)
    got = bb.cuda()(pts.cuda())
    for k in ("sa_indices", "fp_indices"):
        for a, b in zip(got[k], want[k]):
            assert torch.equal(a.cpu(), b), k
    for a, b in zip(got["sa_xyz"], want["sa_xyz"]):
        assert torch.equal(a.cpu(), b)
    for a, b in zip(got["fp_features"], want["fp_features"]):
        assert rel_err(a.detach().cpu(), b.detach()) < 2e-5
    assert got["fp_features"][-1].shape == (2, 256, 256)
    assert got["fp_indices"][-1].dtype == torch.int64


def test_backbone_with_precomputed_fps_chain_is_identical():
    """forward(points, fps_indices=fps_chain(points)) -- the input-pipeline entry bench.py uses --
    returns exactly what forward(points) returns; the after_level hook fires once per SA level."""
    torch.manual_seed(3)
    pts = make_batch(2, 8192, seed0=13)[0].cuda()
    bb = nb.PointNet2SASSG(in_channels=4, num_points=(512, 256, 128, 64), radius=(0.2, 0.4, 0.8, 1.2),
                           num_samples=(32, 16, 16, 16)).cuda()
    bb2 = copy.deepcopy(bb)
    want = bb(pts)
    chain = bb2.fps_chain(pts)
    assert [tuple(c.shape) for c in chain] == [(2, 512), (2, 256), (2, 128), (2, 64)]
    # the chain in two pieces (first level now, the rest later) is the same chain
    first = bb2.fps_chain(pts, stop=1)
    rest = bb2.fps_chain(pts, given=first)
    assert len(first) == 1 and len(rest) == 3
    for a, b in zip(first + rest, chain):
        assert torch.equal(a, b)
    seen = []
    got = bb2(pts, fps_indices=chain, after_level=seen.append)
    assert seen == [0, 1, 2, 3]
    for k in ("sa_indices", "fp_indices", "sa_xyz"):
        for a, b in zip(got[k], want[k]):
            assert torch.equal(a, b), k
    for a, b in zip(got["fp_features"], want["fp_features"]):
        assert torch.equal(a, b)


def test_state_dict_names_follow_reference_layout():
    bb = nb.PointNet2SASSG(in_channels=4)
    keys = set(bb.state_dict().keys())
    assert "SA_modules.0.mlps.0.layer0.conv.weight" in keys
    assert "SA_modules.3.mlps.0.layer2.bn.running_var" in keys
    assert "FP_modules.1.mlps.layer1.bn.weight" in keys
    assert bb.SA_modules[1].mlps[0].layer0.conv.weight.shape == (128, 131, 1, 1)
    assert bb.FP_modules[0].mlps.layer0.conv.weight.shape == (256, 512, 1, 1)
    assert sum(p.numel() for p in bb.parameters()) > 600000
