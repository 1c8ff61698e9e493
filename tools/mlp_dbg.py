import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from nesie_b200 import mlp_rows
from test_mlp_rows_gpu import _layers, _reference
rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
for R, chs, pool_k, xg in [(262144, (4, 64, 64, 128), 64, False), (262144, (4, 64, 64, 128), 64, True), (32768, (4, 64, 64, 128), 64, False), (262144, (4, 64), 0, False), (262144, (8, 64, 64), 0, False)]:
    layers = _layers(chs, torch.float32, 1); ref = _layers(chs, torch.float64, 1)
    torch.manual_seed(5)
    x = (torch.randn(R, chs[0], device="cuda") + 0.5).requires_grad_(xg)
    xd = x.detach().double().requires_grad_(xg)
    got = mlp_rows.mlp_rows(x, layers, pool_k); want = _reference(xd, ref, pool_k)
    g = torch.randn_like(got); got.backward(g); want.backward(g.double())
    print(R, chs, pool_k, xg, "out", f"{rel(got, want):.1e}", " ".join(
        f"L{i}: w {rel(w.grad, wd.grad):.1e} g {rel(bn.weight.grad, bnd.weight.grad):.1e} b {rel(bn.bias.grad, bnd.bias.grad):.1e}"
        for i, ((w, bn), (wd, bnd)) in enumerate(zip(layers, ref))), flush=True)

print("---- unfused fp32 path (linear_rows + bn_relu_rows) vs float64")
from nesie_b200.linear_rows import linear_rows
from nesie_b200 import bn_rows
for R, chs, pool_k in [(262144, (4, 64, 64, 128), 64), (262144, (8, 64, 64), 0)]:
    layers = _layers(chs, torch.float32, 1); ref = _layers(chs, torch.float64, 1)
    torch.manual_seed(5)
    x = (torch.randn(R, chs[0], device="cuda") + 0.5)
    xd = x.detach().double()
    h = x
    for i, (w, bn) in enumerate(layers):
        h = bn_rows.bn_relu_rows(linear_rows(h, w), bn, pool_k if i == len(layers) - 1 else 0)
    want = _reference(xd, ref, pool_k)
    g = torch.randn_like(h); h.backward(g); want.backward(g.double())
    print(R, chs, pool_k, "out", f"{rel(h, want):.1e}", " ".join(
        f"L{i}: w {rel(w.grad, wd.grad):.1e} g {rel(bn.weight.grad, bnd.weight.grad):.1e} b {rel(bn.bias.grad, bnd.bias.grad):.1e}"
        for i, ((w, bn), (wd, bnd)) in enumerate(zip(layers, ref))), flush=True)
