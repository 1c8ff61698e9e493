"""Records outputs of the REFERENCE's own kernels (oracle/_ref, compiled unmodified from
/root/reference/mmdet3d/ops/*/src/*_cuda.cu for sm_100a) into ref_kernels_golden.npz.

Runs on the GPU box (needs a GPU, does not need /root/reference: the .so is prebuilt):
    gpurun -- python tests/golden/make_golden_gpu.py gpurun_out/ref_kernels_golden.npz
The file is then committed under tests/golden/ and pins the C restatement on the CPU
(tests/test_oracle_cpu.py::test_c_oracle_matches_reference_kernels).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nesie_b200.synthetic import make_batch  # noqa: E402
from oracle import ref_cuda  # noqa: E402

# (B, N, M, nsample, min_r, max_r, C, kind)
CASES = [
    (2, 1500, 64, 16, 0.0, 0.3, 5, "scene"),
    (1, 4096, 128, 32, 0.0, 0.4, 4, "scene"),
    (1, 1000, 100, 8, 0.1, 0.5, 3, "scene"),
    (2, 37, 37, 4, 0.0, 0.8, 2, "scene"),
    (1, 3, 3, 2, 0.0, 1.0, 1, "scene"),
    (1, 2049, 200, 8, 0.0, 1.5, 2, "grid"),
    (1, 10000, 256, 16, 0.0, 0.2, 3, "scene"),
    (1, 1, 1, 3, 0.0, 0.2, 1, "scene"),
    (3, 600, 600, 5, 0.0, 1.1, 2, "grid"),
]


def main(out_path):
    assert ref_cuda.available(), "needs oracle/_ref/libnesie_ref_ops.so and a GPU"
    out = {"n_cases": np.int64(len(CASES))}
    rng = np.random.default_rng(7)
    for i, (B, N, M, K, r0, r1, C, kind) in enumerate(CASES):
        if kind == "scene":
            xyz = make_batch(B, max(N, 8), seed0=100 + i)[0][:, :N, :3].contiguous()
        else:  # small-integer grid: exact arithmetic, heavy duplicates / ties
            xyz = torch.from_numpy(rng.integers(0, 4, (B, N, 3)).astype(np.float32))
        feats = torch.from_numpy(rng.standard_normal((B, C, N)).astype(np.float32))
        w = rng.uniform(0.05, 1, (B, N, 3)).astype(np.float32)
        w = torch.from_numpy(w / w.sum(-1, keepdims=True))
        xg, fg, wg = xyz.cuda(), feats.cuda(), w.cuda()
        idx = ref_cuda.furthest_point_sample(xg, M)
        centres = torch.gather(xg, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        bq = ref_cuda.ball_query(r0, r1, K, xg, centres)
        grouped = ref_cuda.grouping_operation(fg, bq)
        gathered = ref_cuda.gather_points(fg, idx)
        dist, i3 = ref_cuda.three_nn_dist2(xg, centres)
        interp = ref_cuda.three_interpolate(gathered, i3, wg)
        torch.cuda.synchronize()
        out.update({f"c{i}_xyz": xyz.numpy(), f"c{i}_m": np.int64(M), f"c{i}_nsample": np.int64(K),
                    f"c{i}_min_r": np.float64(r0), f"c{i}_max_r": np.float64(r1),
                    f"c{i}_feats": feats.numpy(), f"c{i}_weight": w.numpy(),
                    f"c{i}_fps": idx.cpu().numpy(), f"c{i}_bq": bq.cpu().numpy(),
                    f"c{i}_grouped": grouped.cpu().numpy(), f"c{i}_gathered": gathered.cpu().numpy(),
                    f"c{i}_nn_dist2": dist.cpu().numpy(), f"c{i}_nn_idx": i3.cpu().numpy(),
                    f"c{i}_interp": interp.cpu().numpy()})
    pts = torch.from_numpy(rng.standard_normal((1, 300, 6)).astype(np.float32))
    d = ((pts[:, :, None, :] - pts[:, None, :, :]) ** 2).sum(-1).contiguous()
    out["fd_dist"], out["fd_m"] = d.numpy(), np.int64(40)
    out["fd_idx"] = ref_cuda.furthest_point_sample_with_dist(d.cuda(), 40).cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else
         os.path.join(ROOT, "tests", "golden", "ref_kernels_golden.npz"))
