"""Generates tests/golden/sidepool_golden.npz from the REFERENCE's own SidePooling class.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_golden_sidepool.py
`import mmdet3d` is impossible here, so the classes SidePooling / MiniPointNet and rot_gpu are lifted
out of mmdet3d/models/dense_heads/side_pooling_module.py with `ast` and exec'd unmodified on the CPU;
`mmcv.ops.three_nn` is supplied by the C oracle (bit-identical indices to the reference kernel, see
tests/test_oracle_cpu.py) and `Tensor.cuda()` is a no-op.  Only inputs, outputs and a parameter
checksum are stored: parameters are re-created from the seed (same construction order), which also
checks that nesie_b200.SidePooling keeps the reference's state_dict layout."""
import ast
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cpu  # noqa: E402

REF = "/root/reference/mmdet3d/models/dense_heads/side_pooling_module.py"
OUT = os.path.join(HERE, "sidepool_golden.npz")
SEED = 20261018


def lift():
    tree = ast.parse(open(REF).read())
    ns = {"torch": torch, "np": np, "nn": torch.nn, "F": torch.nn.functional,
          "three_nn": lambda target, source: cpu.three_nn(target, source)}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)):
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
    return ns["SidePooling"]


def make_inputs(g, B, K, N, C):
    center = torch.rand(B, K, 3, generator=g) * 4 - 2
    size = torch.rand(B, K, 3, generator=g) * 1.5 + 0.2
    heading = (torch.rand(B, K, generator=g) - 0.5) * 1.0
    seeds = torch.rand(B, N, 3, generator=g) * 5 - 2.5
    feats = torch.randn(B, C, N, generator=g)
    probs = torch.softmax(torch.randn(B, 6, 33, K // 2, generator=g), dim=2)
    return center, size, heading, seeds, feats, probs


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self           # the reference calls .cuda() on grids
    RefSidePooling = lift()
    tmp = os.path.join(tempfile.mkdtemp(), "mean.npz")
    np.savez(tmp, np.ones((4, 3), dtype=np.float32))
    out = {"seed": np.int64(SEED)}
    for case, (B, K, N, C, ncls) in enumerate([(2, 12, 96, 13, 3), (1, 6, 40, 5, 1)]):
        torch.manual_seed(SEED + case)
        ref = RefSidePooling(ncls, 1, ncls, tmp, K // 2, "vote", seed_feat_dim=C)
        g = torch.Generator().manual_seed(SEED + 100 + case)
        center, size, heading, seeds, feats, probs = make_inputs(g, B, K, N, C)
        out[f"c{case}_shape"] = np.array([B, K, N, C, ncls], dtype=np.int64)
        for name, t in [("center", center), ("size", size), ("heading", heading), ("seeds", seeds),
                        ("feats", feats), ("probs", probs)]:
            out[f"c{case}_{name}"] = t.numpy()
        out[f"c{case}_param_abs_sum"] = np.float64(sum(p.detach().double().abs().sum()
                                                       for p in ref.parameters()))
        out[f"c{case}_keys"] = np.array(sorted(ref.state_dict().keys()))
        for mode in ("train", "eval"):
            ref.train(mode == "train")
            ep = {"seed_points": seeds, "seed_features": feats, "bbox_probs": probs}
            with torch.no_grad():
                ep = ref(center, size, heading, ep)
            out[f"c{case}_{mode}_side_scores"] = ep["side_scores"].numpy()
            out[f"c{case}_{mode}_iou_scores"] = ep["iou_scores"].numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
