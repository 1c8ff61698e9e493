"""Records outputs of the REFERENCE's points_in_boxes kernels (oracle/_ref/libnesie_ref_pib.so =
mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu compiled unmodified for sm_100a) on a B200:
    gpurun -- 'python tests/golden/make_golden_pib_gpu.py'   ->  gpurun_out/pib_golden.npz
(copied to tests/golden/pib_golden.npz).  Inputs + reference outputs only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_cuda  # noqa: E402


def case(seed, B, M, T, yaw):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, M, 3, generator=g) * 8 - 4
    pts[..., 2] = torch.rand(B, M, generator=g) * 3
    ctr = torch.rand(B, T, 3, generator=g) * 6 - 3
    ctr[..., 2] = torch.rand(B, T, generator=g) * 1.5
    size = torch.rand(B, T, 3, generator=g) * 1.8 + 0.2
    rz = (torch.rand(B, T, 1, generator=g) - 0.5) * 6.0 if yaw else torch.zeros(B, T, 1)
    boxes = torch.cat([ctr, size, rz], dim=-1)
    # points exactly on faces / centres / corners of the first boxes (the strict '<' cases)
    k = min(T, M // 4)
    pts[:, :k] = ctr[:, :k] + torch.stack([size[:, :k, 1] * 0, size[:, :k, 0] * 0, size[:, :k, 2] / 2], -1)
    if not yaw:
        pts[:, k:2 * k, 0] = ctr[:, :k, 0] + size[:, :k, 0] / 2     # yaw 0: local_x ~ y, local_y ~ -x
        pts[:, k:2 * k, 1] = ctr[:, :k, 1]
        pts[:, k:2 * k, 2] = ctr[:, :k, 2] + 0.1
    return pts, boxes


def main():
    assert ref_cuda.pib_available(), "needs a GPU and oracle/_ref/libnesie_ref_pib.so"
    out = {}
    cases = [(1, 2, 4096, 24, False), (2, 2, 4096, 24, True), (3, 1, 1000, 1, False), (4, 3, 257, 70, True),
             (5, 1, 33, 5, False)]
    for i, (seed, B, M, T, yaw) in enumerate(cases):
        pts, boxes = case(seed, B, M, T, yaw)
        out[f"c{i}_pts"], out[f"c{i}_boxes"] = pts.numpy(), boxes.numpy()
        out[f"c{i}_first"] = ref_cuda.points_in_boxes_gpu(pts.cuda(), boxes.cuda()).cpu().numpy()
        out[f"c{i}_batch"] = ref_cuda.points_in_boxes_batch(pts.cuda(), boxes.cuda()).cpu().numpy().astype(np.int8)
        out[f"c{i}_yaw"] = np.int64(yaw)
    out["count"] = np.int64(len(cases))
    dst = os.path.join(ROOT, "gpurun_out", "pib_golden.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes; inside fraction",
          float(np.mean([out[f"c{i}_batch"].mean() for i in range(len(cases))])))


if __name__ == "__main__":
    main()
