"""Fused training-mode BatchNorm + ReLU (+ group max-pool) on row-major activations vs torch."""
import pytest
import torch
from torch.nn import functional as F

from nesie_b200 import bn_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,C,K", [(4096, 64, 0), (100000, 128, 0), (8192, 256, 16), (65536, 128, 64),
                                   (1024, 128, 32), (50, 12, 0), (2048, 256, 0)])
def test_bn_relu_rows_matches_torch(R, C, K):
    torch.manual_seed(R + C + K)
    y = (torch.randn(R, C, device="cuda") * (1 + torch.rand(C, device="cuda") * 3)
         + torch.randn(C, device="cuda") * 5)
    bn = torch.nn.BatchNorm1d(C).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
    ref = torch.nn.BatchNorm1d(C).cuda()
    ref.load_state_dict(bn.state_dict())
    y1 = y.clone().requires_grad_(True)
    y2 = y.clone().requires_grad_(True)
    assert bn_rows.supported(y1, bn, K)
    got = bn_rows.bn_relu_rows(y1, bn, K)
    want = F.relu(ref(y2))
    if K:
        want = F.max_pool1d(want.view(R // K, K, C).transpose(1, 2), K).squeeze(-1)
    scale = want.abs().max()
    assert ((got - want).abs().max() / scale).item() < 1e-5
    g = torch.randn_like(want)
    got.backward(g)
    want.backward(g)
    gs = y2.grad.abs().max()
    assert ((y1.grad - y2.grad).abs().max() / gs).item() < 2e-5
    assert torch.allclose(bn.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-4 * ref.weight.grad.abs().max().item())
    assert torch.allclose(bn.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-4 * ref.bias.grad.abs().max().item())
    assert torch.allclose(bn.running_mean, ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var, ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == 1
