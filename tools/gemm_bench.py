"""3xTF32 tcgen05 GEMM vs torch.matmul (cuBLAS strict fp32) at the SA shared-MLP shapes, batch 8."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesie_b200.linear_rows import gemm_nt  # noqa: E402
torch.backends.cuda.matmul.allow_tf32 = False

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

for name, R, N, K in [("SA1.l1 fwd", 1048576, 64, 4), ("SA1.l2 fwd", 1048576, 64, 64), ("SA1.l3 fwd", 1048576, 128, 64),
                      ("SA1.l3 dgrad", 1048576, 64, 128), ("SA2.l1 fwd", 262144, 128, 131), ("SA2.l2 fwd", 262144, 128, 128),
                      ("SA2.l3 fwd", 262144, 256, 128), ("SA2.l3 dgrad", 262144, 128, 256), ("SA3.l1 fwd", 65536, 128, 259)]:
    a = torch.randn(R, K, device="cuda"); w = torch.randn(N, K, device="cuda")
    t_mine = timeit(lambda: gemm_nt(a, w)); t_ref = timeit(lambda: a @ w.t())
    want = a[:4096].double() @ w.double().t()
    e_mine = ((gemm_nt(a[:4096].contiguous(), w).double() - want).abs().max() / want.abs().max()).item()
    e_ref = (((a[:4096] @ w.t()).double() - want).abs().max() / want.abs().max()).item()
    fl = 2.0 * R * N * K
    print(json.dumps({"gemm": name, "R": R, "N": N, "K": K, "ms": round(t_mine, 4), "cublas_fp32_ms": round(t_ref, 4),
                      "speedup": round(t_ref / t_mine, 2), "fp32_TFLOPs": round(fl / t_mine / 1e9, 1),
                      "GBps": round((R * K + R * N) * 4 / t_mine / 1e6, 0), "err": f"{e_mine:.1e}", "cublas_err": f"{e_ref:.1e}"}), flush=True)
from nesie_b200.linear_rows import wgrad  # noqa: E402
for name, R, N, K in [("SA1.l2 wgrad", 1048576, 64, 64), ("SA1.l3 wgrad", 1048576, 128, 64), ("SA2.l1 wgrad", 262144, 128, 131),
                      ("SA2.l3 wgrad", 262144, 256, 128), ("SA3.l1 wgrad", 65536, 128, 259)]:
    gy = torch.randn(R, N, device="cuda"); x = torch.randn(R, K, device="cuda")
    t_mine = timeit(lambda: wgrad(gy, x)); t_ref = timeit(lambda: gy.t() @ x)
    want = gy.double().t() @ x.double()
    e_mine = ((wgrad(gy, x).double() - want).abs().max() / want.abs().max()).item()
    e_ref = (((gy.t() @ x).double() - want).abs().max() / want.abs().max()).item()
    print(json.dumps({"gemm": name, "R": R, "N": N, "K": K, "ms": round(t_mine, 4), "cublas_fp32_ms": round(t_ref, 4),
                      "speedup": round(t_ref / t_mine, 2), "GBps": round((R * K + R * N) * 4 / t_mine / 1e6, 0),
                      "err": f"{e_mine:.1e}", "cublas_err": f"{e_ref:.1e}"}), flush=True)
