"""MiniPointNet forward + backward, pooled GEMM epilogues (pool_rows.py) vs the step-by-step kernels:
CUDA events, 256 MB L2 flush between iterations, per-kernel breakdown with NESIE_POOL_PROFILE=1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200.side_pooling import SidePooling  # noqa: E402


def run(sp, mpn, rows, G, iters=12):
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    params = list(mpn.parameters())
    ts = []
    for it in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sp._mini_pointnet(mpn, rows, G)
        torch.autograd.grad(out.sum(), params)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    torch.manual_seed(0)
    sp = SidePooling(18, 1, 18, None, 128, "vote", seed_feat_dim=256).cuda()
    for G, boxes in ((16, 4096), (64, 4096)):
        rows = torch.randn(boxes * G, 260, device="cuda")
        mpn = sp.mlps_before[0]
        for fuse in ("1", "0"):
            os.environ["NESIE_POOL_FUSE"] = fuse
            ms = run(sp, mpn, rows, G)
            print(f"G={G} rows={boxes * G} NESIE_POOL_FUSE={fuse}: {ms * 1000:.1f} us fwd+bwd")
            if os.environ.get("NESIE_POOL_PROFILE") == "1":
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for _ in range(3):
                        out = sp._mini_pointnet(mpn, rows, G)
                        torch.autograd.grad(out.sum(), list(mpn.parameters()))
                    torch.cuda.synchronize()
                for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
                    print(f"    {ev.device_time_total / 3:8.1f} us x{ev.count // 3:3d}  {ev.key[:110]}")


if __name__ == "__main__":
    main()
