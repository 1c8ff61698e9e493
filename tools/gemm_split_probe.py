"""Does a 256-column row GEMM run faster as two 128-column launches (3 pipeline stages instead of 2,
half the weight re-streaming per tile, A read twice)?  Times both for the step's dominant shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesie_b200 import _lib  # noqa: E402
from nesie_b200 import linear_rows as lr  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for R, K, N in [(65536, 256, 256), (65536, 128, 256), (65536, 260, 256), (262144, 256, 256),
                (262144, 128, 256), (65536, 256, 128), (1048576, 64, 128)]:
    ncopy = 6
    xs = [torch.randn(R, K, device=dev) for _ in range(ncopy)]
    ys = [torch.empty(R, N, device=dev) for _ in range(ncopy)]
    w = torch.randn(N, K, device=dev)
    img = lr._pack(w, N, K, K, 1)
    it = [0]

    def whole():
        i = it[0] % ncopy
        it[0] += 1
        _lib.call("nesie_gemm_nt_3xtf32", R, N, K, _lib.ptr(xs[i]), K, _lib.ptr(img), _lib.ptr(ys[i]), N,
                  _lib.stream())
    res = {"whole": timeit(whole)}
    for nb in (128, 64):
        if N % nb or nb >= N:
            continue
        imgs = [lr._pack(w[n0:n0 + nb].contiguous(), nb, K, K, 1) for n0 in range(0, N, nb)]

        def split():
            i = it[0] % ncopy
            it[0] += 1
            for j, im in enumerate(imgs):
                _lib.call("nesie_gemm_nt_3xtf32", R, nb, K, _lib.ptr(xs[i]), K, _lib.ptr(im),
                          _lib.ptr(ys[i]) + 4 * j * nb, N, _lib.stream())
        res[f"split{nb}"] = timeit(split)
    ref = xs[0] @ w.t()
    it[0] = 0
    split() if N > 128 else whole()
    torch.cuda.synchronize()
    err = float((ys[0] - ref).abs().max() / ref.abs().max())
    print(f"R={R} K={K} N={N}: " + ", ".join(f"{k} {v:.1f} us" for k, v in res.items()) + f"  (rel err {err:.1e})")
    del xs, ys
