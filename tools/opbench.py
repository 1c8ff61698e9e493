"""Per-op microbenchmark at the BASELINE shapes (config[1]: batch 8, 40k-point scenes):
this repo's kernels vs the reference's own kernels compiled for sm_100a (oracle/_ref), CUDA-event
timed, L2 flushed between iterations.  Prints one JSON line per op.

    python tools/opbench.py [--batch 8] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesie_b200 as nb  # noqa: E402
from nesie_b200.synthetic import make_batch  # noqa: E402
from oracle import ref_cuda  # noqa: E402

FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    FLUSH.zero_()


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--points", type=int, default=40000)
    args = ap.parse_args()
    B = args.batch
    pts = make_batch(B, args.points, seed0=0)[0].cuda()
    xyz = pts[..., :3].contiguous()
    feat = pts[..., 3:].transpose(1, 2).contiguous()
    have_ref = ref_cuda.available()
    rows = []

    def report(name, mine, ref, alg_bytes=None, note=""):
        med, best = timeit(mine, args.iters)
        row = {"op": name, "ms": round(med, 4), "ms_best": round(best, 4)}
        if alg_bytes:
            row["alg_GBps"] = round(alg_bytes / med / 1e6, 1)
        if have_ref and ref is not None:
            rmed, _ = timeit(ref, max(3, args.iters // 4))
            row["ref_ms"] = round(rmed, 4)
            row["speedup_vs_ref_kernel"] = round(rmed / med, 2)
        if note:
            row["note"] = note
        print(json.dumps(row), flush=True)
        rows.append(row)

    levels = [(args.points, 2048, 0.2, 64, 1), (2048, 1024, 0.4, 32, 128), (1024, 512, 0.8, 16, 256),
              (512, 256, 1.2, 16, 256)]
    cur_xyz = xyz
    cur_feat = feat
    for li, (N, M, r, K, C) in enumerate(levels):
        x = cur_xyz
        report(f"fps_{N}to{M}", lambda: nb.furthest_point_sample(x, M),
               lambda: ref_cuda.furthest_point_sample(x, M), B * (12 * N + 4 * M))
        idx = nb.furthest_point_sample(x, M)
        centres = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        os.environ["NESIE_BALL_QUERY"] = "brute"
        report(f"ball_query_brute_{N}x{M}_k{K}", lambda: nb.ball_query(0.0, r, K, x, centres),
               lambda: ref_cuda.ball_query(0.0, r, K, x, centres), B * (12 * N + 12 * M + 4 * M * K))
        os.environ["NESIE_BALL_QUERY"] = "grid"
        report(f"ball_query_grid_{N}x{M}_k{K}", lambda: nb.ball_query(0.0, r, K, x, centres),
               None, B * (12 * N + 12 * M + 4 * M * K), "build + query")
        os.environ["NESIE_BALL_QUERY"] = "auto"
        bq = nb.ball_query(0.0, r, K, x, centres)
        f = cur_feat if li == 0 else torch.randn(B, C, N, device="cuda")
        report(f"group_points_c{C}_{M}x{K}", lambda: nb.grouping_operation(f, bq),
               lambda: ref_cuda.grouping_operation(f, bq), B * (4 * M * K + 8 * C * M * K))
        grouper = nb.QueryAndGroup(r, K, use_xyz=True, normalize_xyz=True)
        report(f"query_group_concat_c{C}+3_{M}x{K}",
               lambda: nb.group_points._QueryGroupConcat.apply(x, centres, f, bq, r), None,
               B * (4 * M * K + 8 * (C + 3) * M * K), "fused xyz-sub-div-concat")
        cur_xyz = centres
    # FP layers
    for (n, m, C) in [(512, 256, 256), (1024, 512, 256)]:
        t = torch.rand(B, n, 3, device="cuda")
        s = torch.rand(B, m, 3, device="cuda")
        report(f"three_nn_{n}x{m}", lambda: nb.three_nn(t, s), lambda: ref_cuda.three_nn(t, s),
               B * (12 * n + 12 * m + 24 * n))
        _, i3 = nb.three_nn(t, s)
        w = torch.rand(B, n, 3, device="cuda")
        f = torch.randn(B, C, m, device="cuda")
        report(f"three_interpolate_c{C}_{m}to{n}", lambda: nb.three_interpolate(f, i3, w),
               lambda: ref_cuda.three_interpolate(f, i3, w), B * (16 * C * n + 24 * n))
    return rows


if __name__ == "__main__":
    main()
