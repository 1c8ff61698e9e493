"""Reads a chrome trace written by `NESIE_BENCH_TRACE=path python bench.py ...` and reports, for the
middle step, how long a large kernel (GEMM / BatchNorm / FPS / row kernels) was active, how long only
small kernels ran, and the longest stretches of the latter with the kernels they contain."""
import collections
import json
import sys

t = json.load(open(sys.argv[1]))
ev = sorted((e for e in t["traceEvents"] if e.get("cat") == "kernel"), key=lambda e: e["ts"])
ad = [e for e in ev if "FusedOptimizer" in e["name"]]
w0, w1 = ad[0]["ts"] + ad[0]["dur"], ad[1]["ts"] + ad[1]["dur"]
win = [e for e in ev if w0 <= e["ts"] < w1]
BIG = ("gemm_", "bn_relu", "bn_colsum", "fps_reg", "interp_rows", "three_nn", "group_max", "group_rows_kernel")


def big(e):
    return any(k in e["name"] for k in BIG)


pts = []
for e in win:
    pts += [(e["ts"], 1, big(e)), (e["ts"] + e["dur"], -1, big(e))]
pts.sort()
nb = ns = 0
last, tb, tsml, tidle, start, stretches = w0, 0.0, 0.0, 0.0, w0, []
for ts, d, b in pts:
    dt = ts - last
    if nb > 0:
        tb += dt
    elif ns > 0:
        tsml += dt
    else:
        tidle += dt
    last = ts
    if b:
        if nb == 0 and d == 1 and start is not None:
            stretches.append((start, ts))
            start = None
        nb += d
        if nb == 0:
            start = ts
    else:
        ns += d
print(f"step {1e-3 * (w1 - w0):.2f} ms, {len(win)} kernels: large kernel active {tb / 1e3:.2f} ms, "
      f"only small kernels {tsml / 1e3:.2f} ms, idle {tidle / 1e3:.2f} ms")
for a, b in sorted(stretches, key=lambda s: s[0] - s[1])[:8]:
    names = collections.Counter(e["name"][:48] for e in win if a <= e["ts"] < b)
    print(f"  at {(a - w0) / 1e3:6.2f} ms, {b - a:5.0f} us, {sum(names.values())} kernels: {names.most_common(3)}")
