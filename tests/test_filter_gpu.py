"""Parity of the pseudo-label filter pieces on the GPU: aligned_3d_nms, the lenient float64 NMS,
the device-side get_pseudo_labels, the fused side-uncertainty loss (forward + backward) and the
flat-buffer teacher EMA -- against vectors from the reference's own source (py_golden.npz) and
against the oracle restatements on fresh inputs.  Keep-lists are bit-exact."""
import numpy as np
import pytest
import torch

import nesie_b200 as nb
from oracle import restate

pytestmark = pytest.mark.gpu


def test_aligned_3d_nms_golden(golden_py):
    g = golden_py
    for i in range(int(g["aligned_count"])):
        keep = nb.aligned_3d_nms(torch.from_numpy(g[f"aligned_{i}_boxes"]).cuda(),
                                 torch.from_numpy(g[f"aligned_{i}_scores"]).cuda(),
                                 torch.from_numpy(g[f"aligned_{i}_classes"]).cuda(),
                                 float(g[f"aligned_{i}_thr"]))
        assert keep.dtype == torch.int64
        assert np.array_equal(keep.cpu().numpy(), g[f"aligned_{i}_keep"]), i


def _rand_boxes(rng, n):
    anchors = rng.uniform(-4, 4, (max(1, n // 5), 3))
    c = anchors[rng.integers(0, len(anchors), n)] + rng.normal(0, 0.2, (n, 3))
    s = rng.uniform(0.1, 1.5, (n, 3))
    return np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)


def test_aligned_3d_nms_batched_vs_oracle_and_idempotence():
    rng = np.random.default_rng(11)
    S, n = 16, 256
    boxes = np.stack([_rand_boxes(rng, n) for _ in range(S)])
    scores = np.stack([rng.permutation(n).astype(np.float32) / n for _ in range(S)])
    classes = rng.integers(0, 18, (S, n))
    counts = rng.integers(1, n + 1, S).astype(np.int32)
    counts[0], counts[1] = n, 1
    keep, cnt = nb.aligned_3d_nms_batched(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                                          torch.from_numpy(classes).cuda(), 0.25,
                                          torch.from_numpy(counts).cuda())
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    for s in range(S):
        c = counts[s]
        want = restate.aligned_3d_nms(boxes[s, :c], scores[s, :c], classes[s, :c], 0.25)
        assert np.array_equal(keep[s, :cnt[s]], want), s
        assert (keep[s, cnt[s]:] == -1).all()
        # idempotence: NMS of the kept set keeps everything, in the same order
        again = nb.aligned_3d_nms(torch.from_numpy(boxes[s][want]).cuda(), torch.from_numpy(scores[s][want]).cuda(),
                                  torch.from_numpy(classes[s][want]).cuda(), 0.25)
        assert again.tolist() == list(range(len(want)))


def test_aligned_3d_nms_empty():
    keep = nb.aligned_3d_nms(torch.zeros(0, 6).cuda(), torch.zeros(0).cuda(),
                             torch.zeros(0, dtype=torch.long).cuda(), 0.25)
    assert keep.numel() == 0


def test_lhs_nms_golden(golden_py):
    g = golden_py
    for i in range(int(g["lhs_count"])):
        rows = torch.from_numpy(g[f"lhs_{i}_rows"]).cuda()
        pick, cnt = nb.lhs_3d_faster_samecls_batched(rows[None], float(g[f"lhs_{i}_thr"]),
                                                     bool(g[f"lhs_{i}_old"]))
        assert pick[0, :int(cnt[0])].tolist() == g[f"lhs_{i}_pick"].tolist(), i


def test_pseudo_label_filter_matches_restatement():
    torch.manual_seed(3)
    B, P, C = 4, 256, 18
    for trial in range(3):
        bbox = torch.cat([torch.randn(B, P, 3) * 1.5, torch.rand(B, P, 3) * 1.5 + 0.2,
                          torch.zeros(B, P, 1)], -1)
        # cluster proposals so that the NMS has something to suppress
        bbox[:, :, :3] = bbox[:, torch.randint(0, 12, (P,)), :3] + 0.1 * torch.randn(B, P, 3)
        preds = dict(bbox_preds=bbox, sem_scores=torch.rand(B, P, C) ** 0.3,
                     obj_scores=torch.randn(B, P, 2) * 5, iou_scores=torch.rand(B, P, C),
                     side_scores=torch.rand(B, P, 6, C), vote_points=torch.rand(B, P, 3))
        ulb_list = torch.randint(0, 6, (40, C)).float()
        ulb_flag = (torch.rand(40) > 0.5).float()
        want = restate.get_pseudo_labels(preds, ulb_list, ulb_flag, 12, 40)
        got = nb.get_pseudo_labels({k: v.clone().cuda() for k, v in preds.items()}, ulb_list.cuda(),
                                   ulb_flag.cuda(), 12, 40)
        n_total = 0
        for b in range(B):
            assert torch.equal(got[0][b].cpu().long(), want[0][b].long()), (trial, b)
            assert torch.equal(got[1][b].cpu(), want[1][b]), (trial, b)
            assert torch.allclose(got[2][b].cpu(), want[2][b], rtol=1e-6, atol=1e-6), (trial, b)
            n_total += want[0][b].shape[0]
        assert n_total > 0  # the case is not vacuous


def test_side_uncertainty_loss_forward_backward():
    torch.manual_seed(5)
    rows, C = 8 * 256, 18
    pred = torch.randn(rows, 6)
    tgt = torch.cat([torch.randn(rows, 3), torch.rand(rows, 3) + 0.2, torch.zeros(rows, 1)], -1)
    side = torch.rand(rows, 6, C)
    sem = torch.randn(rows, C)
    w = (torch.rand(rows, 1) > 0.7).float().repeat(1, 6) / 50.0
    pc, sc = pred.clone().requires_grad_(True), side.clone().requires_grad_(True)
    want, wsig = restate.side_uncertainty_loss(pc, tgt, sc, sem, w, 10.0, 1.0)
    (want + (wsig.mean(-1) * w[:, 0]).sum()).backward()
    pg, sg = pred.cuda().requires_grad_(True), side.cuda().requires_grad_(True)
    got, gsig = nb.side_uncertainty_loss(pg, tgt.cuda(), sg, sem.cuda(), w.cuda(), 10.0, 1.0)
    (got + (gsig.mean(-1) * w[:, 0].cuda()).sum()).backward()
    assert torch.allclose(got.cpu(), want.detach(), rtol=1e-5)
    assert torch.allclose(gsig.detach().cpu(), wsig.detach(), rtol=1e-6, atol=1e-7)
    assert torch.allclose(pg.grad.cpu(), pc.grad, rtol=1e-5, atol=1e-8)
    assert torch.allclose(sg.grad.cpu(), sc.grad, rtol=1e-5, atol=1e-8)


def test_teacher_ema_matches_hook_arithmetic():
    torch.manual_seed(9)
    model = torch.nn.Sequential(torch.nn.Conv1d(7, 33, 1), torch.nn.BatchNorm1d(33),
                                torch.nn.Conv1d(33, 5, 1)).cuda()
    ref_ema = [p.detach().cpu().clone() for p in model.parameters()]
    ema = nb.TeacherEMA(model, momentum=0.001, warm_up=10)
    n_buffers = sum(1 for _ in model.buffers())
    assert n_buffers == 3  # running stats are NOT part of the EMA
    for step in range(5):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        ema.after_train_iter(step)
        for e, p in zip(ref_ema, model.parameters()):
            restate.ema_update(e, p.detach().cpu(), 0.001, 10, step)
    for (name, got), want in zip(ema.ema_state_dict().items(), ref_ema):
        assert name.startswith("ema_")
        assert torch.allclose(got.cpu(), want, rtol=1e-6, atol=1e-7)
    before = [p.detach().clone() for p in model.parameters()]
    ema.swap()
    for p, want in zip(model.parameters(), ref_ema):
        assert torch.allclose(p.detach().cpu(), want, rtol=1e-6, atol=1e-7)
    ema.swap()
    for p, b in zip(model.parameters(), before):
        assert torch.equal(p.detach(), b)
