"""Generates tests/golden/py_golden.npz from the REFERENCE's own python source.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden_py.py
`import mmdet3d` is impossible here (mmcv/mmdet are not installed), so the pure torch / numpy
function bodies are lifted out of the reference files with `ast` and exec'd unmodified:
    aligned_3d_nms          mmdet3d/core/post_processing/box3d_nms.py:129-176
    lhs_3d_faster_samecls   mmdet3d/models/detectors/votenet_nesie.py:733-779
    flip_axis_to_camera, get_3d_box, roty              votenet_nesie.py:781-821
    Bbox2Surface            mmdet3d/models/losses/surface_loss.py:90-100
Only inputs and the reference's outputs are stored; no reference code is copied.
"""
import ast
import os

import numpy as np
import torch

REF = "/root/reference/mmdet3d"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "py_golden.npz")


def lift(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def rand_boxes(rng, n, room=4.0, smin=0.1, smax=1.5, cluster=True):
    c = rng.uniform(-room, room, size=(n, 3)).astype(np.float32)
    if cluster:  # overlapping clusters so that suppression actually happens
        anchors = rng.uniform(-room, room, size=(max(1, n // 6), 3)).astype(np.float32)
        c = anchors[rng.integers(0, len(anchors), n)] + rng.normal(0, 0.15, (n, 3)).astype(np.float32)
    s = rng.uniform(smin, smax, size=(n, 3)).astype(np.float32)
    return np.concatenate([c - s / 2, c + s / 2], axis=1).astype(np.float32), c, s


def main():
    (aligned,) = lift(f"{REF}/core/post_processing/box3d_nms.py", ["aligned_3d_nms"])
    lhs, flip, get_3d_box, roty = lift(f"{REF}/models/detectors/votenet_nesie.py",
                                       ["lhs_3d_faster_samecls", "flip_axis_to_camera",
                                        "get_3d_box", "roty"])
    get_3d_box.__globals__["roty"] = roty
    (b2s,) = lift(f"{REF}/models/losses/surface_loss.py", ["Bbox2Surface"])
    rng = np.random.default_rng(20261018)
    out = {}

    # --- aligned_3d_nms -------------------------------------------------------------------
    cases = []
    for n, ncls, thr in [(256, 18, 0.25), (256, 3, 0.25), (100, 1, 0.5), (1, 18, 0.25),
                         (37, 18, 0.1), (256, 18, 0.0)]:
        boxes, _, _ = rand_boxes(rng, n)
        scores = rng.permutation(n).astype(np.float32) / n + rng.uniform(0, 1e-3)
        classes = rng.integers(0, ncls, n)
        cases.append((boxes, scores.astype(np.float32), classes, thr))
    # degenerate: zero-volume duplicates (NaN IoU drops boxes even across classes)
    boxes, _, _ = rand_boxes(rng, 24)
    boxes[5, 3:] = boxes[5, :3]
    boxes[9] = boxes[5]
    boxes[11, 3] = boxes[11, 0]
    scores = (rng.permutation(24).astype(np.float32) + 1) / 25
    cases.append((boxes, scores, rng.integers(0, 4, 24), 0.25))
    for i, (boxes, scores, classes, thr) in enumerate(cases):
        assert len(np.unique(scores)) == len(scores)
        keep = aligned(torch.from_numpy(boxes), torch.from_numpy(scores),
                       torch.from_numpy(classes), thr).numpy()
        out[f"aligned_{i}_boxes"], out[f"aligned_{i}_scores"] = boxes, scores
        out[f"aligned_{i}_classes"], out[f"aligned_{i}_thr"] = classes.astype(np.int64), np.float64(thr)
        out[f"aligned_{i}_keep"] = keep.astype(np.int64)
    out["aligned_count"] = np.int64(len(cases))

    # --- lhs_3d_faster_samecls ------------------------------------------------------------
    k = 0
    for n, ncls, thr, old in [(64, 18, 0.25, False), (64, 2, 0.25, False), (64, 1, 0.25, True),
                              (7, 1, 0.1, False), (1, 1, 0.25, False), (200, 4, 0.25, False)]:
        boxes, _, _ = rand_boxes(rng, n)
        rows = np.zeros((n, 8))
        rows[:, :6] = boxes
        rows[:, 6] = (rng.permutation(n).astype(np.float32) + 1) / (n + 1)
        rows[:, 7] = rng.integers(0, ncls, n)
        pick = lhs(rows.copy(), thr, old)
        out[f"lhs_{k}_rows"], out[f"lhs_{k}_thr"] = rows, np.float64(thr)
        out[f"lhs_{k}_old"], out[f"lhs_{k}_pick"] = np.int64(old), np.asarray(pick, dtype=np.int64)
        k += 1
    out["lhs_count"] = np.int64(k)

    # --- get_3d_box corner min/max as get_pseudo_labels stores them (:219-250) -------------
    n = 200
    center = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    size = rng.uniform(0.05, 3, (n, 3)).astype(np.float32)
    cam = flip(center)
    corners = np.zeros((n, 8, 3), dtype=np.float32)
    for j in range(n):
        corners[j] = get_3d_box(size[j], np.float32(0.0), cam[j, :])
    mm = np.zeros((n, 6))
    mm[:, :3], mm[:, 3:] = corners.min(axis=1), corners.max(axis=1)
    out["box_center"], out["box_size"], out["box_minmax"] = center, size, mm

    # --- Bbox2Surface ---------------------------------------------------------------------
    b7 = rng.normal(0, 2, (64, 7)).astype(np.float32)
    out["b2s_in"], out["b2s_out"] = b7, b2s(torch.from_numpy(b7)).numpy()

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if "keep" in k or "pick" in k})


if __name__ == "__main__":
    main()
