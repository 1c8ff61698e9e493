"""SidePooling on this repo's kernels (three_nn + nesie_interp_rows + tcgen05 row GEMMs) against the CPU
oracle that is pinned to the reference class (tests/test_side_pooling_cpu.py)."""
import copy
import os

import numpy as np
import pytest
import torch

import nesie_b200 as nb
from nesie_b200 import _lib
from nesie_b200.side_pooling import SidePooling
from oracle.side_pooling_ref import SidePoolingOracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sidepool_golden.npz")


def _inputs(B, K, N, C, seed):
    g = torch.Generator().manual_seed(seed)
    center = torch.rand(B, K, 3, generator=g) * 4 - 2
    size = torch.rand(B, K, 3, generator=g) * 1.5 + 0.2
    heading = (torch.rand(B, K, generator=g) - 0.5) * 1.0
    seeds = torch.rand(B, N, 3, generator=g) * 5 - 2.5
    feats = torch.randn(B, C, N, generator=g)
    probs = torch.softmax(torch.randn(B, 6, 33, K // 2, generator=g), dim=2)
    return (center, size, heading), {"seed_points": seeds, "seed_features": feats, "bbox_probs": probs}


def test_interp_rows_kernel_matches_three_interpolate():
    torch.manual_seed(0)
    B, C, m, n = 2, 37, 200, 999
    xyz = torch.rand(B, m, 3, device="cuda")
    tgt = torch.rand(B, n, 3, device="cuda")
    feats = torch.randn(B, C, m, device="cuda")
    dist, idx = nb.three_nn(tgt, xyz)
    w = 1.0 / (dist + 1e-8)
    w = (w / w.sum(2, keepdim=True)).contiguous()
    want = nb.three_interpolate(feats, idx, w)                       # (B, C, n), reference contraction
    head = torch.randn(B, n, 3, device="cuda")
    ld = 40
    rows = torch.full((B * n, ld), 7.0, device="cuda")
    _lib.call("nesie_interp_rows", B, C, m, n, _lib.ptr(feats.transpose(1, 2).contiguous()), _lib.ptr(idx),
              _lib.ptr(w), _lib.ptr(head), _lib.ptr(rows), ld, _lib.stream())
    rows = rows.view(B, n, ld)
    assert torch.equal(rows[..., :3], head)
    assert torch.equal(rows[..., 3:3 + C], want.transpose(1, 2))     # bit-identical
    assert (rows[..., 3 + C:] == 0).all()


@pytest.mark.parametrize("B,K,N,C,ncls", [(2, 12, 96, 13, 3), (2, 32, 256, 256, 18)])
def test_forward_and_parameter_gradients_match_oracle(B, K, N, C, ncls):
    torch.manual_seed(4)
    ref = SidePoolingOracle(ncls, 1, ncls, None, K // 2, "vote", seed_feat_dim=C)
    gpu = SidePooling(ncls, 1, ncls, None, K // 2, "vote", seed_feat_dim=C)
    gpu.load_state_dict(copy.deepcopy(ref.state_dict()))
    gpu = gpu.cuda()
    boxes, ep = _inputs(B, K, N, C, 11)
    want = ref(*boxes, dict(ep))
    got = gpu(*[t.cuda() for t in boxes], {k: v.cuda() for k, v in ep.items()})
    for key in ("side_scores", "iou_scores"):
        scale = want[key].abs().max().clamp_min(1.0)
        assert (got[key].cpu() - want[key]).abs().max() < 1e-4 * scale, key
    gs = torch.randn_like(want["side_scores"])
    gi = torch.randn_like(want["iou_scores"])
    ((want["side_scores"] * gs).sum() + (want["iou_scores"] * gi).sum()).backward()
    ((got["side_scores"] * gs.cuda()).sum() + (got["iou_scores"] * gi.cuda()).sum()).backward()
    pg = dict(gpu.named_parameters())
    gmax = max(float(p.grad.norm()) for p in ref.parameters())
    for name, p in ref.named_parameters():
        g = pg[name].grad.cpu()
        err = float((g - p.grad).norm())
        # L2 (ReLU-boundary flips, see test_votenet_gpu.py).  Biases in front of a training-mode
        # BatchNorm have a mathematically zero gradient: only rounding noise on both sides.
        assert err < 2e-2 * float(p.grad.norm()) or err < 1e-5 * gmax, (name, err, float(p.grad.norm()))
    # running statistics moved identically
    for (n1, b1), (n2, b2) in zip(ref.named_buffers(), gpu.named_buffers()):
        assert n1 == n2
        assert torch.allclose(b1.float(), b2.cpu().float(), rtol=1e-4, atol=1e-5), n1
    # eval mode (running statistics)
    ref.eval(); gpu.eval()
    with torch.no_grad():
        want = ref(*boxes, dict(ep))
        got = gpu(*[t.cuda() for t in boxes], {k: v.cuda() for k, v in ep.items()})
    for key in ("side_scores", "iou_scores"):
        scale = want[key].abs().max().clamp_min(1.0)
        assert (got[key].cpu() - want[key]).abs().max() < 1e-4 * scale, key


def test_gpu_module_against_reference_golden():
    gold = np.load(GOLD)
    case = 0
    B, K, N, C, ncls = (int(v) for v in gold[f"c{case}_shape"])
    torch.manual_seed(int(gold["seed"]) + case)
    mod = SidePooling(ncls, 1, ncls, None, K // 2, "vote", seed_feat_dim=C).cuda()
    t = lambda name: torch.from_numpy(gold[f"c{case}_{name}"]).cuda()  # noqa: E731
    ep = {"seed_points": t("seeds"), "seed_features": t("feats"), "bbox_probs": t("probs")}
    with torch.no_grad():
        out = mod(t("center"), t("size"), t("heading"), ep)
    for key in ("side_scores", "iou_scores"):
        want = torch.from_numpy(gold[f"c{case}_train_{key}"])
        assert (out[key].cpu() - want).abs().max() < 1e-4 * want.abs().max().clamp_min(1.0), key


def test_refuses_cpu_tensors():
    mod = SidePooling(3, 1, 3, None, 4, "vote", seed_feat_dim=5)
    boxes, ep = _inputs(1, 8, 20, 5, 1)
    with pytest.raises((RuntimeError, AssertionError)):
        mod(*boxes, ep)


@pytest.mark.parametrize("k,concat", [(16, False), (16, True), (64, True), (5, False)])
def test_group_max_rows_matches_torch(k, concat):
    from nesie_b200.group_max import group_max_concat_rows, group_max_rows
    torch.manual_seed(k)
    groups, C = 300, 128
    x = torch.randn(groups * k, C, device="cuda", requires_grad=True)
    b = torch.randn(C, device="cuda", requires_grad=True)
    xr, br = x.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    f = (xr + br).view(groups, k, C)
    gmax = f.max(dim=1).values
    want = torch.cat([gmax.unsqueeze(1).expand(-1, k, -1), f], dim=2).reshape(groups * k, 2 * C) if concat else gmax
    got = (group_max_concat_rows if concat else group_max_rows)(x, b, k)
    assert torch.equal(got, want)
    g = torch.randn_like(want)
    got.backward(g)
    want.backward(g)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-5, atol=2e-5)   # sum over k rows: order differs
    assert torch.allclose(b.grad, br.grad, rtol=1e-4, atol=1e-4)
