"""Fused tcgen05 SA forward vs the unfused eval-mode path (gather kernels + cuDNN fp32), BASELINE
shapes, batch 8.  Prints one JSON line per SA level with achieved TFLOP/s of the fused kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesie_b200 as nb  # noqa: E402
from nesie_b200 import sa_fused  # noqa: E402
from nesie_b200.synthetic import make_batch  # noqa: E402

LEVELS = [  # name, N, M, K, r, C, mlp
    ("SA1", 40000, 2048, 64, 0.2, 1, [64, 64, 128]),
    ("SA2", 2048, 1024, 32, 0.4, 128, [128, 128, 256]),
    ("SA3", 1024, 512, 16, 0.8, 256, [128, 128, 256]),
    ("SA4", 512, 256, 16, 1.2, 256, [128, 128, 256]),
    ("AGG", 1024, 256, 16, 0.3, 256, [128, 128, 128]),
]


def timeit(fn, iters=20):
    for _ in range(3):
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
    B = 8
    only = sys.argv[1] if len(sys.argv) > 1 else None
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    xyz_all = make_batch(B, 40000, seed0=0)[0][..., :3].contiguous().cuda()
    for name, N, M, K, r, C, mlp in LEVELS:
        if only and name != only:
            continue
        xyz = xyz_all[:, :N].contiguous()
        feats = torch.randn(B, C, N, device="cuda")
        sa = nb.PointSAModule(mlp_channels=[C] + mlp, num_point=M, radius=r, num_sample=K,
                              use_xyz=True, normalize_xyz=True).cuda().eval()
        idx_fps = nb.furthest_point_sample(xyz, M)
        centres = nb.gather_points(xyz.transpose(1, 2).contiguous(), idx_fps).transpose(1, 2).contiguous()
        bq = nb.ball_query(0.0, r, K, xyz, centres)
        packed = sa_fused.fold_mlp(sa.mlps[0], C)
        chans = [C + 3] + mlp
        flops = 2.0 * B * M * K * sum(a * b for a, b in zip(chans[:-1], chans[1:]))
        with torch.no_grad():
            def unfused():
                g = sa.groupers[0](xyz, centres, feats)
                return sa._pool_features(sa.mlps[0](g))

            def fused():
                return sa_fused.sa_fused_forward(xyz, centres, feats, bq, r, packed)

            table = sa_fused.pack_features(feats, B, N)
            out = torch.empty((B, mlp[-1], M), device="cuda")
            from nesie_b200 import _lib

            def kernel_only():
                _lib.call("nesie_sa_fused_forward", B, N, M, K, C, mlp[0], mlp[1], mlp[2],
                          _lib.ptr(xyz), _lib.ptr(centres), _lib.ptr(table), _lib.ptr(bq), float(r),
                          _lib.ptr(packed['w1']), _lib.ptr(packed['w2']), _lib.ptr(packed['w3']),
                          _lib.ptr(packed['scale_shift']), _lib.ptr(out), _lib.stream())

            a, b = unfused(), fused()
            err = ((a - b).abs().max() / a.abs().max()).item()
            t_un, t_fu, t_k = timeit(unfused), timeit(fused), timeit(kernel_only)
        print(json.dumps({"level": name, "rows": B * M * K, "gflop": round(flops / 1e9, 2),
                          "unfused_ms": round(t_un, 4), "fused_ms": round(t_fu, 4),
                          "fused_kernel_ms": round(t_k, 4),
                          "kernel_tflops": round(flops / t_k / 1e9, 1),
                          "speedup": round(t_un / t_fu, 2), "rel_err": round(err, 5)}), flush=True)


if __name__ == "__main__":
    main()
