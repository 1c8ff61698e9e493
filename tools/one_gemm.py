import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200.linear_rows import gemm_nt
R, N, K = [int(v) for v in sys.argv[1:4]]
a = torch.randn(R, K, device="cuda"); w = torch.randn(N, K, device="cuda")
out = gemm_nt(a, w); torch.cuda.synchronize()
want = a.double() @ w.double().t()
print("err", ((out.double() - want).abs().max() / want.abs().max()).item())
