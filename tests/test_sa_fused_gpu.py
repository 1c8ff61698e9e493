"""Fused tcgen05 SA forward (eval mode, folded BN, bf16 operands) vs the unfused fp32 path of the
same module (this repo's gather kernels + cuDNN fp32): the north star's "bf16 MLP within 1e-2"."""
import pytest
import torch

import nesie_b200 as nb
from nesie_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu

SHAPES = [  # N, M, K, radius, C_in, mlp
    (8192, 512, 64, 0.2, 1, [64, 64, 128]),      # SA1 shape (fewer points)
    (2048, 1024, 32, 0.4, 128, [128, 128, 256]),  # SA2
    (1024, 512, 16, 0.8, 256, [128, 128, 256]),   # SA3
    (512, 256, 16, 1.2, 256, [128, 128, 256]),    # SA4
    (1024, 256, 16, 0.3, 256, [128, 128, 128]),   # vote aggregation
]


@pytest.mark.parametrize("N,M,K,r,C,mlp", SHAPES)
def test_fused_matches_unfused(N, M, K, r, C, mlp):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(N + C)
    B = 3
    xyz = make_batch(B, max(N, 8), seed0=60)[0][:, :N, :3].contiguous().cuda()
    feats = torch.randn(B, C, N, device="cuda")
    sa = nb.PointSAModule(mlp_channels=[C] + mlp, num_point=M, radius=r, num_sample=K,
                          use_xyz=True, normalize_xyz=True).cuda()
    with torch.no_grad():  # non-trivial BN statistics / affine parameters
        for layer in sa.mlps[0]:
            layer.bn.running_mean.normal_(0, 0.3)
            layer.bn.running_var.uniform_(0.5, 1.5)
            layer.bn.weight.uniform_(0.5, 1.5)
            layer.bn.bias.normal_(0, 0.2)
    sa.eval()
    with torch.no_grad():
        x0, f0, i0 = sa(xyz, feats)
        sa.fused_bf16 = True
        assert sa._fused_ok(0, feats)
        x1, f1, i1 = sa(xyz, feats)
    assert torch.equal(i0, i1) and torch.equal(x0, x1)
    err = ((f1 - f0).abs().max() / f0.abs().max()).item()
    assert err < 1e-2, err
    # elementwise: bf16-level agreement almost everywhere
    close = torch.isclose(f1, f0, rtol=3e-2, atol=3e-2 * f0.abs().max().item())
    assert close.float().mean().item() > 0.999
