"""Times nesie_fps at the BASELINE shape under the tuning knobs (cluster size, threads per CTA,
exchange mechanism) and checks every variant against the default's indices."""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesie_b200 as nb  # noqa: E402
from nesie_b200.synthetic import make_batch  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    xyz = make_batch(B, N, seed0=0)[0][..., :3].contiguous().cuda()
    for k in ("NESIE_FPS_CLUSTER", "NESIE_FPS_THREADS", "NESIE_FPS_XMODE"):
        os.environ.pop(k, None)
    base = nb.furthest_point_sample(xyz, M)
    for cl, nt, xm in itertools.product([4, 8, 16], [128, 256], [0, 1]):
        os.environ["NESIE_FPS_CLUSTER"] = str(cl)
        os.environ["NESIE_FPS_THREADS"] = str(nt)
        os.environ["NESIE_FPS_XMODE"] = str(xm)
        try:
            got = nb.furthest_point_sample(xyz, M)
            ok = bool(torch.equal(got, base))
            ms = timeit(lambda: nb.furthest_point_sample(xyz, M))
            print(json.dumps({"B": B, "N": N, "M": M, "cluster": cl, "threads": nt, "xmode": xm,
                              "ms": round(ms, 4), "us_per_iter": round(ms * 1000 / (M - 1), 3),
                              "same_as_default": ok}), flush=True)
        except RuntimeError as e:
            print(json.dumps({"cluster": cl, "threads": nt, "xmode": xm, "error": str(e)[:120]}))


if __name__ == "__main__":
    main()
