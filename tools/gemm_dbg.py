"""Bottleneck experiments on the NT GEMM (NESIE_GEMM_DBG bits: 1 no A loads, 2 no C stores, 4 no MMAs)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200.linear_rows import gemm_nt  # noqa: E402

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

for R, N, K in [(1048576, 64, 64), (1048576, 128, 64), (262144, 256, 128), (262144, 128, 256)]:
    a = torch.randn(R, K, device="cuda"); w = torch.randn(N, K, device="cuda")
    out = []
    for dbg in [0]:
        os.environ["NESIE_GEMM_DBG"] = str(dbg)
        out.append(f"dbg{dbg}={timeit(lambda: gemm_nt(a, w)) * 1e3:.0f}us")
        os.environ["NESIE_GEMM_DBG"] = str(dbg | 128)
        import ctypes
        from nesie_b200 import _lib
        buf = (ctypes.c_longlong * 16)()
        _lib.lib().nesie_gemm_debug_profile(buf)
        gemm_nt(a, w)
        _lib.lib().nesie_gemm_debug_profile(buf)
        v = list(buf)
        ns = max(v[3], 1); nt = max(v[10], 1)
        out.append(f"[loader/slab: wait {v[0]//ns} fill {v[1]//ns} publish {v[2]//ns} (n={v[3]}) | mma/tile: acce {v[4]//nt} full {v[5]//nt} issue {v[6]//nt} total {v[7]//nt} | epi/tile: wait {v[8]//nt} drain {v[9]//nt} (n={v[10]})]")
    print(R, N, K, " ".join(out), flush=True)

from nesie_b200.linear_rows import wgrad  # noqa: E402
import ctypes  # noqa: E402
from nesie_b200 import _lib  # noqa: E402
for R, N, K in [(1048576, 64, 64), (1048576, 128, 64), (1048576, 64, 128), (262144, 256, 128), (262144, 128, 256), (262144, 128, 128)]:
    gy = torch.randn(R, N, device="cuda"); x = torch.randn(R, K, device="cuda")
    want = gy.double().t() @ x.double()
    for layout in ["3", "2", "0"]:
        os.environ["NESIE_WGRAD_LAYOUT"] = layout
        os.environ["NESIE_GEMM_DBG"] = "0"
        t = timeit(lambda: wgrad(gy, x)) * 1e3
        err = ((wgrad(gy, x).double() - want).abs().max() / want.abs().max()).item()
        os.environ["NESIE_GEMM_DBG"] = "128"
        buf = (ctypes.c_longlong * 16)()
        _lib.lib().nesie_gemm_debug_profile(buf)
        wgrad(gy, x)
        _lib.lib().nesie_gemm_debug_profile(buf)
        v = list(buf); ns = max(v[3], 1); nt = max(v[10], 1)
        print("wgrad", R, N, K, "layout", layout, f"{t:.0f}us err {err:.1e} [loader/slab: wait {v[0]//ns} fill {v[1]//ns} publish {v[2]//ns} | mma/chunk: full {v[5]//nt} issue {v[6]//nt} total {v[7]//nt} (chunks {v[10]}, slabs {v[3]})]", flush=True)
