"""On-disk format and sampling: `.bin` round trip, the height channel against the
reference's own LoadPointsFromFile lines lifted from its source, IndoorPointSample's two regimes, and
the checkpoint key split."""
import ast
import os

import numpy as np
import pytest
import torch

from nesie_b200 import data

REF = "/root/reference/mmdet3d/datasets/pipelines/loading.py"


def _scene(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n, 6)).astype(np.float32)


def test_bin_round_trip_and_height_channel(tmp_path):
    raw = _scene(5000)
    path = os.path.join(tmp_path, "scene.bin")
    data.save_points(path, raw)
    assert os.path.getsize(path) == raw.size * 4
    pts = data.load_points(path)
    assert pts.dtype == np.float32 and pts.shape == (5000, 4)
    assert np.array_equal(pts[:, :3], raw[:, :3])
    floor = np.percentile(raw[:, 2], 0.99)
    assert np.array_equal(pts[:, 3], raw[:, 2] - floor)
    assert np.array_equal(data.points_from_bytes(raw.tobytes()), pts)
    np.save(os.path.join(tmp_path, "scene.npy"), raw)
    assert np.array_equal(data.load_points(os.path.join(tmp_path, "scene.npy")), pts)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")
def test_height_channel_matches_reference_source():
    """The statements of LoadPointsFromFile.__call__ (loading.py:411-420) exec'd from the reference."""
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LoadPointsFromFile")
    call = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__call__")
    body = [s for s in call.body if not isinstance(s, ast.Expr)][:6]     # up to the shift_height block
    src = ast.Module([ast.FunctionDef("f", call.args, body + [ast.Return(ast.Name("points", ast.Load()))],
                                      [], lineno=1, col_offset=0)], [])
    ast.fix_missing_locations(src)
    ns = {"np": np}
    exec(compile(src, REF, "exec"), ns)
    raw = _scene(3000, 1)

    class Self:
        load_dim, use_dim, shift_height = 6, [0, 1, 2], True

        @staticmethod
        def _load_points(name):
            return raw.reshape(-1)
    want = ns["f"](Self, dict(pts_filename="x"))
    assert np.array_equal(data.points_from_bytes(raw.tobytes()), want)


def test_indoor_point_sample_regimes():
    g = torch.Generator().manual_seed(0)
    big = torch.arange(100000, dtype=torch.float32).unsqueeze(1).repeat(1, 4)
    p, c = data.indoor_point_sample(big, 40000, g)
    assert p.shape == (40000, 4) and c.unique().numel() == 40000          # without replacement
    assert torch.equal(p[:, 0].long(), c)
    small = big[:1000]
    p, c = data.indoor_point_sample(small, 4000, g)
    assert p.shape == (4000, 4) and int(c.max()) < 1000 and c.unique().numel() < 4000   # with replacement
    p2, _ = data.indoor_point_sample(small, 4000, choices=c)
    assert torch.equal(p, p2)


def test_scene_batcher_and_views(tmp_path):
    files = []
    for i in range(4):
        f = os.path.join(tmp_path, f"s{i}.bin")
        data.save_points(f, _scene(3000 + 500 * i, i))
        files.append(f)
    b = data.SceneBatcher(files, num_points=2048, batch_size=2, device="cpu")
    batches = list(b)
    assert len(batches) == 2 and batches[0][0].shape == (2, 2048, 4)
    pts, choices = batches[1]
    assert torch.equal(pts[0], torch.from_numpy(data.load_points(files[2]))[choices[0]])
    boxes = torch.rand(2, 5, 7)
    v = data.mean_teacher_views(pts, boxes, torch.Generator().manual_seed(1))
    assert v["points_s"].shape == pts.shape and v["gt_boxes_s"].shape == boxes.shape
    # heights ride along unchanged, the two views differ
    assert torch.equal(v["points_s"][..., 3], pts[..., 3]) and not torch.equal(v["points_s"], v["points_t"])


def test_split_reference_state_dict():
    state = {"backbone.SA_modules.0.mlps.0.layer0.conv.weight": torch.zeros(1),
             "ema_backbone_SA_modules_0_mlps_0_layer0_conv_weight": torch.ones(1)}
    w, e = data.split_reference_state_dict(state)
    assert list(w) == ["backbone.SA_modules.0.mlps.0.layer0.conv.weight"]
    assert list(e) == ["ema_backbone_SA_modules_0_mlps_0_layer0_conv_weight"]
