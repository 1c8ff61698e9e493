"""Pins the oracle (test infrastructure) on the CPU:
  * python-level restatements vs vectors produced by the reference's own source
    (tests/golden/py_golden.npz, generator: tests/golden/make_golden_py.py);
  * the C restatement vs outputs of the reference's own kernels recorded on a B200
    (tests/golden/ref_kernels_golden.npz, generator: tests/golden/make_golden_gpu.py);
  * the C restatement vs independent brute-force numpy on small cases.
"""
import numpy as np
import pytest
import torch

from oracle import cpu, restate


def test_aligned_3d_nms_matches_reference_source(golden_py):
    g = golden_py
    for i in range(int(g["aligned_count"])):
        keep = restate.aligned_3d_nms(g[f"aligned_{i}_boxes"], g[f"aligned_{i}_scores"],
                                      g[f"aligned_{i}_classes"], float(g[f"aligned_{i}_thr"]))
        assert np.array_equal(keep, g[f"aligned_{i}_keep"]), i


def test_lhs_nms_matches_reference_source(golden_py):
    g = golden_py
    for i in range(int(g["lhs_count"])):
        pick = restate.lhs_3d_faster_samecls(g[f"lhs_{i}_rows"], float(g[f"lhs_{i}_thr"]),
                                             bool(g[f"lhs_{i}_old"]))
        assert pick == g[f"lhs_{i}_pick"].tolist(), i


def test_corner_minmax_matches_reference_source(golden_py):
    g = golden_py
    c, s = g["box_center"], g["box_size"]
    for j in range(c.shape[0]):
        cam = np.array([c[j, 0], -c[j, 2], c[j, 1]], dtype=np.float32)
        lo, hi = restate._get_3d_box_minmax(s[j], cam)
        assert np.array_equal(np.concatenate([lo, hi]).astype(np.float64), g["box_minmax"][j])


def test_bbox2surface_matches_reference_source(golden_py):
    out = restate.bbox2surface(torch.from_numpy(golden_py["b2s_in"])).numpy()
    assert np.array_equal(out, golden_py["b2s_out"])


def test_c_oracle_matches_reference_kernels(golden_ref):
    """Every case recorded from oracle/_ref on the B200 is reproduced bit-for-bit."""
    g = golden_ref
    n = int(g["n_cases"])
    assert n >= 6
    for i in range(n):
        xyz = torch.from_numpy(g[f"c{i}_xyz"])
        m, k = int(g[f"c{i}_m"]), int(g[f"c{i}_nsample"])
        r0, r1 = float(g[f"c{i}_min_r"]), float(g[f"c{i}_max_r"])
        idx = cpu.furthest_point_sample(xyz, m)
        assert torch.equal(idx, torch.from_numpy(g[f"c{i}_fps"])), f"fps case {i}"
        centres = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        bq = cpu.ball_query(r0, r1, k, xyz, centres)
        assert torch.equal(bq, torch.from_numpy(g[f"c{i}_bq"])), f"ball_query case {i}"
        feats = torch.from_numpy(g[f"c{i}_feats"])
        grouped = cpu.grouping_operation(feats, bq)
        assert torch.equal(grouped, torch.from_numpy(g[f"c{i}_grouped"])), f"group case {i}"
        gathered = cpu.gather_points(feats, idx)
        assert torch.equal(gathered, torch.from_numpy(g[f"c{i}_gathered"])), f"gather case {i}"
        dist2, i3 = cpu.three_nn_dist2(xyz, centres)
        assert torch.equal(i3, torch.from_numpy(g[f"c{i}_nn_idx"])), f"three_nn idx case {i}"
        assert torch.equal(dist2, torch.from_numpy(g[f"c{i}_nn_dist2"])), f"three_nn dist case {i}"
        w = torch.from_numpy(g[f"c{i}_weight"])
        interp = cpu.three_interpolate(gathered, i3, w)
        assert torch.equal(interp, torch.from_numpy(g[f"c{i}_interp"])), f"interpolate case {i}"
    dist = torch.from_numpy(g["fd_dist"])
    assert torch.equal(cpu.furthest_point_sample_with_dist(dist, int(g["fd_m"])),
                       torch.from_numpy(g["fd_idx"]))


def _naive_ball_query(xyz, centres, r0, r1, k):
    out = np.zeros((centres.shape[0], k), dtype=np.int32)
    r0s, r1s = np.float32(r0) * np.float32(r0), np.float32(r1) * np.float32(r1)
    for j, c in enumerate(centres):
        d = c[None, :] - xyz
        # exact products on the small-integer grid used below: no rounding anywhere
        d2 = (d * d).sum(1)
        hits = np.where((d2 == 0) | ((d2 >= r0s) & (d2 < r1s)))[0][:k]
        if len(hits):
            out[j, :] = hits[0]
            out[j, :len(hits)] = hits
    return out


def test_ball_query_against_bruteforce_on_exact_grid():
    rng = np.random.default_rng(3)
    xyz = rng.integers(0, 6, (1, 700, 3)).astype(np.float32)
    centres = xyz[:, :50].copy()
    for r0, r1, k in [(0.0, 1.5, 8), (1.0, 2.5, 16), (0.0, 0.5, 4)]:
        got = cpu.ball_query(r0, r1, k, torch.from_numpy(xyz), torch.from_numpy(centres))[0].numpy()
        assert np.array_equal(got, _naive_ball_query(xyz[0], centres[0], r0, r1, k))


def test_ball_query_empty_ball_is_zero_row():
    xyz = torch.zeros(1, 10, 3) + 5.0
    centres = torch.zeros(1, 2, 3)
    assert torch.equal(cpu.ball_query(0.0, 0.1, 4, xyz, centres), torch.zeros(1, 2, 4, dtype=torch.int32))


def test_fps_tie_rule_is_bit_reversed_slot_order():
    """Equal maxima are resolved by the reference's shared-memory tree, i.e. lowest
    (bitreverse(k mod bs), k div bs); exercised on a tiny integer grid full of duplicates."""
    def brev32(x):
        return int('{:032b}'.format(x)[::-1], 2)
    rng = np.random.default_rng(0)
    for n in [5, 64, 100, 513, 1500, 3000]:
        xyz = rng.integers(0, 4, size=(n, 3)).astype(np.float32)
        m = min(n, 60)
        bs = cpu.opt_n_threads(n)
        p = bs.bit_length() - 1
        key = np.array([(brev32(k % bs) if p else 0) | (k // bs) for k in range(n)], dtype=np.uint64)
        temp = np.full(n, 1e10, np.float32)
        want, old = [0], 0
        for _ in range(1, m):
            temp = np.minimum(temp, ((xyz - xyz[old]) ** 2).sum(1).astype(np.float32))
            cand = np.where(temp == temp.max())[0]
            old = int(cand[np.argmin(key[cand])])
            want.append(old)
        got = cpu.furthest_point_sample(torch.from_numpy(xyz)[None], m)[0].numpy()
        assert np.array_equal(got, np.array(want)), n


def test_three_nn_fewer_than_three_sources():
    t = torch.rand(1, 4, 3)
    s = torch.rand(1, 2, 3)
    dist, idx = cpu.three_nn(t, s)
    assert torch.isinf(dist[..., 2]).all() and (idx[..., 2] == 0).all()


def test_grad_restatements_are_adjoint():
    """<gather(f), g> == <f, gather_grad(g)> for the three scatter ops."""
    torch.manual_seed(0)
    f = torch.rand(2, 3, 50)
    idx = torch.randint(0, 50, (2, 7, 4), dtype=torch.int32)
    g = torch.rand(2, 3, 7, 4)
    a = (cpu.grouping_operation(f, idx) * g).sum()
    b = (f * cpu.grouping_operation_grad(g, idx, 50)).sum()
    assert torch.allclose(a, b, rtol=1e-5)
    i3 = torch.randint(0, 50, (2, 9, 3), dtype=torch.int32)
    w = torch.rand(2, 9, 3)
    g2 = torch.rand(2, 3, 9)
    a = (cpu.three_interpolate(f, i3, w) * g2).sum()
    b = (f * cpu.three_interpolate_grad(g2, i3, w, 50)).sum()
    assert torch.allclose(a, b, rtol=1e-5)


def test_pseudo_label_restatement_runs_and_respects_masks():
    torch.manual_seed(1)
    B, P, C = 2, 256, 18
    preds = dict(bbox_preds=torch.rand(B, P, 7) * 2, sem_scores=torch.rand(B, P, C),
                 obj_scores=torch.randn(B, P, 2) * 4, iou_scores=torch.rand(B, P, C),
                 side_scores=torch.rand(B, P, 6, C), vote_points=torch.rand(B, P, 3))
    ulb_list = torch.randint(0, 5, (30, C)).float()
    labels, boxes, quals = restate.get_pseudo_labels(preds, ulb_list, torch.ones(30), 12, 30)
    assert len(labels) == B
    for l, b, q in zip(labels, boxes, quals):
        assert b.shape[0] == l.shape[0] == q.shape[0] and b.shape[0] <= 64
