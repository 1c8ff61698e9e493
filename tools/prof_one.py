"""Runs ONE hot-path op at the BASELINE shape a few times (for `ncu -k regex:<kernel> -s <skip> -c 1`).

    python tools/prof_one.py fps|ball_query|group|interpolate|three_nn|nms [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesie_b200 as nb  # noqa: E402
from nesie_b200.synthetic import make_batch  # noqa: E402


def main():
    op = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    pts = make_batch(B, 40000, seed0=0)[0].cuda()
    xyz = pts[..., :3].contiguous()
    idx = nb.furthest_point_sample(xyz, 2048)
    centres = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    reps = 3
    if op == "fps":
        for _ in range(reps):
            nb.furthest_point_sample(xyz, 2048)
    elif op == "ball_query":
        for _ in range(reps):
            nb.ball_query(0.0, 0.2, 64, xyz, centres)
    elif op == "group":
        f = torch.randn(B, 128, 2048, device="cuda")
        bq = torch.randint(0, 2048, (B, 1024, 32), dtype=torch.int32, device="cuda")
        for _ in range(reps):
            nb.grouping_operation(f, bq)
    elif op == "interpolate":
        f = torch.randn(B, 256, 512, device="cuda")
        i3 = torch.randint(0, 512, (B, 1024, 3), dtype=torch.int32, device="cuda")
        w = torch.rand(B, 1024, 3, device="cuda")
        for _ in range(reps):
            nb.three_interpolate(f, i3, w)
    elif op == "three_nn":
        t, s = torch.rand(B, 1024, 3, device="cuda"), torch.rand(B, 512, 3, device="cuda")
        for _ in range(reps):
            nb.three_nn(t, s)
    torch.cuda.synchronize()
    print("ok", op)


if __name__ == "__main__":
    main()
