import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200.linear_rows import wgrad
R, N, K = 32, 128, 32
for dbg in ["1,256,64", "0,256,64", "1,64,256", "0,1,64", "1,1,64", "1,256,0", "1,0,0"]:
    os.environ["NESIE_WGRAD_DBG"] = dbg
    gy = torch.ones(R, N, device="cuda"); x = torch.ones(R, K, device="cuda")
    got = wgrad(gy, x); torch.cuda.synchronize()
    gy2 = torch.randn(R, N, device="cuda"); x2 = torch.randn(R, K, device="cuda")
    g2 = wgrad(gy2, x2); w2 = gy2.t() @ x2
    print(dbg, "ones ->", got.min().item(), got.max().item(), "| rand err", float((g2 - w2).abs().max()), flush=True)
