import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from nesie_b200 import mlp_rows
from test_mlp_rows_gpu import _layers, _reference
rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
for R, chs, pool_k in [(70000, (64, 64), 0), (4096, (64, 64), 0), (70000, (64, 64, 64), 0), (18944, (64, 64), 0), (19072, (64, 64), 0), (70016, (64, 64), 0)]:
    layers = _layers(chs, torch.float32, 1); ref = _layers(chs, torch.float64, 1)
    torch.manual_seed(5)
    x = (torch.randn(R, chs[0], device="cuda") + 0.5).requires_grad_(True)
    xd = x.detach().double().requires_grad_(True)
    got = mlp_rows.mlp_rows(x, layers, pool_k); want = _reference(xd, ref, pool_k)
    g = torch.randn_like(got); got.backward(g); want.backward(g.double())
    bad = (x.grad.double() - xd.grad).abs().max(dim=1).values
    print(R, chs, "out", rel(got, want), "xgrad", rel(x.grad, xd.grad), "wgrad", rel(layers[0][0].grad, ref[0][0].grad),
          "dgamma", rel(layers[0][1].weight.grad, ref[0][1].weight.grad), "bad rows", (bad > 1e-3).nonzero().flatten()[:6].tolist(), int((bad > 1e-3).sum()))
