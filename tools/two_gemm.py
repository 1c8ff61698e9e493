"""One NT GEMM and one weight-gradient GEMM launch (for an ncu capture): R N K [Rw Nw Kw]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200.linear_rows import gemm_nt, wgrad  # noqa: E402

v = [int(x) for x in sys.argv[1:]]
R, N, K = v[:3]
Rw, Nw, Kw = v[3:6] if len(v) >= 6 else (R, N, K)
a = torch.randn(R, K, device="cuda")
w = torch.randn(N, K, device="cuda")
out = gemm_nt(a, w)
gy = torch.randn(Rw, Nw, device="cuda")
x = torch.randn(Rw, Kw, device="cuda")
g = wgrad(gy, x)
torch.cuda.synchronize()
print("nt err", ((out[:4096].double() - a[:4096].double() @ w.double().t()).abs().max()).item())
print("wgrad err", ((g.double() - gy.double().t() @ x.double()).abs().max() / R ** 0.5).item())
