"""Per-role cycle breakdown of CTA 0 of the weight-gradient kernel (NESIE_GEMM_DBG=128 counters read by
nesie_gemm_debug_profile): transform warps (wait for TMA | work), MMA warp (wait for a free accumulator |
wait for a transformed stage | issue), epilogue (wait | drain).
    NESIE_GEMM_DBG=128 [NESIE_WGRAD_KSPLIT=0|1] python tools/wgrad_profile.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesie_b200 import _lib  # noqa: E402

assert os.environ.get("NESIE_GEMM_DBG") == "128", "run with NESIE_GEMM_DBG=128"
dev = torch.device("cuda:0")
for R, N, K in [(65536, 256, 256), (65536, 128, 256), (262144, 256, 256), (65536, 128, 128), (1048576, 64, 64)]:
    gy = torch.randn(R, N, device=dev)
    x = torch.randn(R, K, device=dev)
    ns = _lib.lib().nesie_gemm_wgrad_splits(R, N, K)
    parts = torch.empty((ns, N, K), device=dev)
    out = (ctypes.c_longlong * 16)()
    for it in range(3):
        _lib.call("nesie_gemm_wgrad_3xtf32", R, N, K, _lib.ptr(gy), N, _lib.ptr(x), K, _lib.ptr(parts), ns,
                  _lib.stream())
        torch.cuda.synchronize()
        _lib.lib().nesie_gemm_debug_profile(out)     # also resets the counters
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.call("nesie_gemm_wgrad_3xtf32", R, N, K, _lib.ptr(gy), N, _lib.ptr(x), K, _lib.ptr(parts), ns,
              _lib.stream())
    b.record()
    torch.cuda.synchronize()
    _lib.lib().nesie_gemm_debug_profile(out)
    v = list(out)
    tot = max(v[7], 1)
    print(f"R={R} N={N} K={K} splits={ns}: {a.elapsed_time(b) * 1e3:.1f} us | CTA0 total {tot} cyc, slabs {v[3]} | "
          f"transform: wait TMA {100 * v[0] / tot:.0f}% work {100 * v[1] / tot:.0f}% | "
          f"MMA warp: wait acc {100 * v[4] / tot:.0f}% wait stage {100 * v[5] / tot:.0f}% issue {100 * v[6] / tot:.0f}% | "
          f"epilogue: wait {v[8]} drain {v[9]} cyc over {v[10]} chunks")
