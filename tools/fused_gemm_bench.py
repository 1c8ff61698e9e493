"""Cost of the fused BatchNorm prologue / column statistics on the NT GEMM (back-to-back launches)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200 import _lib
from nesie_b200 import linear_rows as lr

def run(R, K, N, pro, stats, reps=10):
    a = torch.randn(R, K, device="cuda"); w = torch.randn(N, K, device="cuda"); out = torch.empty(R, N, device="cuda")
    img = lr._pack(w, N, K, K, 1)
    sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    parts = torch.empty(_lib.lib().nesie_gemm_stats_parts(R), 2, N, device="cuda")
    def one():
        _lib.call("nesie_gemm_nt_3xtf32_fused", R, N, K, _lib.ptr(a), K, _lib.ptr(img), _lib.ptr(out), N,
                  _lib.ptr(sc) if pro else None, _lib.ptr(sh) if pro else None, _lib.ptr(parts) if stats else None, _lib.stream())
    for _ in range(3): one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): one()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for R, K, N in [(1048576, 64, 64), (1048576, 64, 128), (1048576, 128, 64), (262144, 128, 128), (262144, 128, 256), (262144, 256, 128), (65536, 128, 256)]:
    t = [run(R, K, N, p, s) for p, s in [(0, 0), (0, 1), (1, 0), (1, 1)]]
    gb = 4.0 * R * (K + N) / 1e3
    print(f"{R:8d} {K:4d}->{N:4d}: plain {t[0]:6.1f} us ({gb / t[0]:.0f} GB/s) | +stats {t[1]:6.1f} | +prologue {t[2]:6.1f} | both {t[3]:6.1f}", flush=True)
