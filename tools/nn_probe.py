"""three_nn brute force vs grid at the SidePooling shape (8 scenes, 1024 seeds, 81920 grid points)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nesie_b200 as nb  # noqa: E402
from nesie_b200.interpolate import three_nn_grid  # noqa: E402
from nesie_b200.synthetic import make_batch  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000


def main():
    torch.manual_seed(0)
    pts = make_batch(8, 40000)[0][..., :3].contiguous().cuda()
    seeds = pts[:, torch.randperm(40000, device="cuda")[:1024]].contiguous()
    for spread in (0.2, 0.6, 2.0, 8.0):
        pick = torch.randint(0, 1024, (8, 81920), device="cuda")
        tgt = (torch.gather(seeds, 1, pick.unsqueeze(-1).expand(-1, -1, 3)) +
               torch.randn(8, 81920, 3, device="cuda") * spread).contiguous()
        t0 = timeit(lambda: nb.three_nn(tgt, seeds))
        t1 = timeit(lambda: three_nn_grid(tgt, seeds))
        print(f"targets within ~{spread} m of a seed: brute {t0:.0f} us, grid {t1:.0f} us")


if __name__ == "__main__":
    main()
