"""Per-role cycle breakdown of CTA 0 of the row GEMM (NESIE_GEMM_DBG=128 counters): MMA warp (wait for the
epilogue to free an accumulator | wait for a transformed stage | issue) and epilogue (wait | drain).
    NESIE_GEMM_DBG=128 python tools/gemm_profile.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesie_b200 import _lib  # noqa: E402
from nesie_b200 import linear_rows as lr  # noqa: E402

assert os.environ.get("NESIE_GEMM_DBG") == "128", "run with NESIE_GEMM_DBG=128"
dev = torch.device("cuda:0")
for R, K, N in [(65536, 256, 256), (65536, 128, 256), (65536, 256, 128), (262144, 256, 256), (65536, 128, 128),
                (1048576, 64, 128), (1048576, 64, 64)]:
    x = torch.randn(R, K, device=dev)
    y = torch.empty(R, N, device=dev)
    img = lr._pack(torch.randn(N, K, device=dev), N, K, K, 1)
    out = (ctypes.c_longlong * 16)()

    def launch():
        _lib.call("nesie_gemm_nt_3xtf32", R, N, K, _lib.ptr(x), K, _lib.ptr(img), _lib.ptr(y), N, _lib.stream())
    for _ in range(3):
        launch()
        torch.cuda.synchronize()
        _lib.lib().nesie_gemm_debug_profile(out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    launch()
    b.record()
    torch.cuda.synchronize()
    _lib.lib().nesie_gemm_debug_profile(out)
    v = list(out)
    tot = max(v[7], 1)
    tiles = max(v[10], 1)
    print(f"R={R} K={K} N={N}: {a.elapsed_time(b) * 1e3:.1f} us | CTA0 {tot} cyc, {tiles} tiles ({tot // tiles} cyc/tile) | "
          f"MMA warp: wait acc {100 * v[4] / tot:.0f}% wait stage {100 * v[5] / tot:.0f}% issue {100 * v[6] / tot:.0f}% | "
          f"transform: wait TMA {100 * v[0] / tot:.0f}% work {100 * v[1] / tot:.0f}% | "
          f"epilogue warp 0: wait {100 * v[8] / tot:.0f}% drain {100 * v[9] / tot:.0f}% ({v[9] // tiles} cyc/tile)")
