"""Times nesie_fps under the tuning knobs (cluster size, threads per CTA) and checks every
variant against the default's indices.  python tools/fps_variants.py [B] [N] [M]"""
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
    for k in ("NESIE_FPS_CLUSTER", "NESIE_FPS_THREADS"):
        os.environ.pop(k, None)
    base = nb.furthest_point_sample(xyz, M)
    ms = timeit(lambda: nb.furthest_point_sample(xyz, M))
    print(json.dumps({"B": B, "N": N, "M": M, "variant": "default", "ms": round(ms, 4),
                      "us_per_iter": round(ms * 1000 / (M - 1), 3)}), flush=True)
    for cl, nt in itertools.product([1, 2, 4, 8, 16], [32, 64, 128, 256]):
        os.environ["NESIE_FPS_CLUSTER"] = str(cl)
        os.environ["NESIE_FPS_THREADS"] = str(nt)
        try:
            got = nb.furthest_point_sample(xyz, M)
            ok = bool(torch.equal(got, base))
            ms = timeit(lambda: nb.furthest_point_sample(xyz, M))
            print(json.dumps({"B": B, "N": N, "M": M, "cluster": cl, "threads": nt,
                              "ms": round(ms, 4), "us_per_iter": round(ms * 1000 / (M - 1), 3),
                              "same_as_default": ok}), flush=True)
        except RuntimeError as e:
            print(json.dumps({"cluster": cl, "threads": nt, "error": str(e)[-90:]}))


if __name__ == "__main__":
    main()
