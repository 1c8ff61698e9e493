"""NesieHead on the GPU kernels against the golden vectors the REFERENCE's own NesieHead source
produced (tests/golden/head_golden.npz): target indices / masks exact, loss terms and gradients
within 1e-5, forward (module re-created from the seed) within 1e-5 of the feature scale."""
import pytest
import torch

from head_cases import (GRAD_KEYS, LOSS_KEYS, TARGET_NAMES, UNSUP_KEYS, load_golden, loss_inputs,
                        make_head, rel_err)

pytestmark = pytest.mark.gpu
G = load_golden()
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


@pytest.fixture(scope="module")
def head():
    from nesie_b200.nesie_head import NesieHead
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(int(G["seed"]))
    return make_head(NesieHead, 16, 8).to(DEV)


@pytest.mark.parametrize("tag", ["lossA", "lossB"])
def test_targets_match_reference(head, tag):
    preds, points, boxes, labels, _ = loss_inputs(G, tag, DEV)
    got = head.get_targets(points, boxes, labels, bbox_preds=preds)
    for name, t in zip(TARGET_NAMES, got):
        if name == "bbox_targets":
            t = torch.cat(t, dim=0)
        want = torch.from_numpy(G[f"{tag}_tgt_{name}"])
        if want.dtype in (torch.int64, torch.int32):
            assert torch.equal(t.long().cpu(), want.long()), name
        else:
            assert torch.allclose(t.cpu(), want, rtol=1e-6, atol=1e-7), name


@pytest.mark.parametrize("tag", ["lossA", "lossB"])
def test_supervised_loss_and_grads_match_reference(head, tag):
    preds, points, boxes, labels, _ = loss_inputs(G, tag, DEV, grad=True)
    losses = head.loss(preds, points, boxes, labels)
    for k in LOSS_KEYS:
        assert rel_err(losses[k], G[f"{tag}_sup_{k}"]) < 1e-5, (k, losses[k].item(), G[f"{tag}_sup_{k}"])
    sum(losses.values()).backward()
    for k in GRAD_KEYS:
        g = preds[k].grad if preds[k].grad is not None else torch.zeros_like(preds[k])
        assert rel_err(g, G[f"{tag}_sup_grad_{k}"]) < 1e-5, k


@pytest.mark.parametrize("tag", ["lossA", "lossB"])
def test_unsupervised_loss_and_grads_match_reference(head, tag):
    preds, points, _, _, pl = loss_inputs(G, tag, DEV, grad=True)
    losses = head.unsup_loss(preds, points, pl[0], pl[1], None, pl[2])
    for k in UNSUP_KEYS:
        assert rel_err(losses[k], G[f"{tag}_unsup_{k}"]) < 1e-5, (k, losses[k].item())
    sum(losses.values()).backward()
    for k in GRAD_KEYS:
        g = preds[k].grad if preds[k].grad is not None else torch.zeros_like(preds[k])
        assert rel_err(g, G[f"{tag}_unsup_grad_{k}"]) < 1e-5, k


def test_vote_targets_at_seeds_equal_gathered_full_targets(head):
    from nesie_b200 import targets as T
    preds, points, boxes, labels, _ = loss_inputs(G, "lossB", DEV)
    bx, lb, valid = T.pad_gt(boxes, labels, DEV, pad_to=16)
    nv = valid.sum(1).int()
    full_t, full_m = T.vote_targets(points, bx, nv)
    seed_t, seed_m = T.vote_targets(points, bx, nv, preds["seed_indices"])
    idx = preds["seed_indices"]
    assert torch.equal(seed_m, torch.gather(full_m, 1, idx))
    assert torch.equal(seed_t, torch.gather(full_t, 1, idx.unsqueeze(-1).expand(-1, -1, 9)))


def test_padding_beyond_the_largest_box_count_changes_nothing(head):
    """loss_padded with GT padded to a static 16 slots == the list interface (padded to max count)."""
    from nesie_b200 import targets as T
    preds, points, boxes, labels, _ = loss_inputs(G, "lossA", DEV)
    a = head.loss(preds, points, boxes, labels)
    b = head.loss_padded(preds, points, *T.pad_gt(boxes, labels, DEV, pad_to=16))
    for k in LOSS_KEYS:   # the padded zeros only change the shape of torch's summation tree
        assert torch.allclose(a[k], b[k], rtol=2e-6, atol=0), k


def test_forward_matches_reference():
    from nesie_b200.nesie_head import NesieHead
    B, S, C, P = (int(v) for v in G["fwd_shape"])
    torch.manual_seed(int(G["seed"]) + 10)
    head = make_head(NesieHead, C, P)
    assert sorted(head.state_dict().keys()) == list(G["fwd_keys"])
    head = head.to(DEV)
    for mode in ("train", "eval"):
        head.train(mode == "train")
        feat = dict(fp_xyz=[torch.from_numpy(G["fwd_seed_points"]).to(DEV)],
                    fp_features=[torch.from_numpy(G["fwd_seed_features"]).to(DEV)],
                    fp_indices=[torch.from_numpy(G["fwd_seed_indices"]).to(DEV)])
        torch.manual_seed(int(G["seed"]) + 12)
        noise = tuple(torch.randn(B, P, 3).to(DEV) for _ in range(2))
        with torch.no_grad():
            res = head(feat, "vote", "ScanNet", jitter_noise=noise)
        for k in ["vote_points", "vote_features", "aggregated_points", "aggregated_features", "obj_scores",
                  "sem_scores", "surface_pred", "bbox_preds", "bbox_probs", "jitter_bbox_preds",
                  "iou_scores", "iou_scores_jitter", "side_scores", "side_scores_jitter"]:
            assert rel_err(res[k], G[f"fwd_{mode}_{k}"]) < 2e-5, (mode, k)
        assert torch.equal(res["aggregated_indices"].long().cpu(),
                           torch.from_numpy(G[f"fwd_{mode}_aggregated_indices"]).long())


def test_sort_vertices_equals_reference_kernel_live():
    """nesie_sort_vertices against the reference's own kernel compiled unmodified (oracle/_ref) on
    random polygons, including vertices exactly on the axes through the mean."""
    from oracle import ref_cuda
    if not ref_cuda.sortv_available():
        pytest.skip("oracle/_ref/libnesie_ref_sortv.so not built")
    from nesie_b200.rotated_iou import sort_vertices
    g = torch.Generator().manual_seed(5)
    v = torch.rand(4, 600, 24, 2, generator=g) - 0.5
    v[:, :50, :, 1] = torch.round(v[:, :50, :, 1] * 4) / 4      # exact zeros and ties
    m = torch.rand(4, 600, 24, generator=g) > 0.7
    m[:, :, 8:][:, :, -1] = False                               # at least one free intersection slot
    nv = m.int().sum(-1).int()
    keep = nv <= 8
    m = m & keep.unsqueeze(-1)
    nv = m.int().sum(-1).int()
    v, m, nv = v.to(DEV), m.to(DEV), nv.to(DEV)
    got = sort_vertices(v, m, nv)
    want = ref_cuda.sort_vertices(v, m, nv)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


def test_fused_iou3d_matches_tensor_formulation():
    """nesie_iou3d (value + gradient in one kernel) against the reference's tensor formulation run on
    the same GPU (itself bit-identical to the reference source on the CPU, tests/test_head_cpu.py):
    rotated, axis-aligned, identical, disjoint and touching boxes."""
    from nesie_b200.rotated_iou import cal_iou_3d, cal_iou_3d_tensor
    g = torch.Generator().manual_seed(11)
    B, N = 4, 700
    a = torch.cat([torch.rand(B, N, 3, generator=g) * 2, torch.rand(B, N, 3, generator=g) + 0.2,
                   torch.randn(B, N, 1, generator=g)], -1)
    b = a.clone()
    b[..., :6] += torch.randn(B, N, 6, generator=g) * 0.25
    b[..., 3:6].clamp_(min=0.05)
    b[..., 6] = 0
    b[:, :30] = a[:, :30]                  # identical boxes
    a[:, 30:80, 6] = 0                     # both axis aligned
    b[:, 80:110, :3] += 10.0               # disjoint
    a, b = a.to(DEV), b.to(DEV)
    a1 = a.clone().requires_grad_(True)
    a2 = a.clone().requires_grad_(True)
    want = cal_iou_3d_tensor(a1, b)
    got = cal_iou_3d(a2, b)
    assert torch.allclose(got, want, rtol=2e-6, atol=2e-7), float((got - want).abs().max())
    w = torch.rand(B, N, generator=g).to(DEV)
    (want * w).sum().backward()
    (got * w).sum().backward()
    scale = float(a1.grad.abs().max())
    assert float((a1.grad - a2.grad).abs().max()) < 2e-5 * scale
    # no gradient requested: forward only
    assert torch.equal(cal_iou_3d(a, b), got.detach())


@pytest.mark.parametrize("per_class", [True, False])
def test_get_bboxes_matches_reference(per_class):
    """Test-time decoding (get_bboxes / multiclass_nms_single, nesie_head.py:681-788) against the
    reference's own source: objectness x IoU score, non-empty test, aligned_3d_nms over the non-empty
    boxes, score threshold, per-class expansion."""
    from head_cases import check_decoded, decode_inputs
    from nesie_b200.nesie_head import NesieHead
    torch.manual_seed(0)
    head = make_head(NesieHead, 16, 8, test_cfg=dict(nms_thr=0.25, score_thr=0.05, per_class_proposal=per_class))
    head = head.to(DEV)
    points, preds = decode_inputs(G, DEV)
    check_decoded(G, per_class, head.get_bboxes(points, preds))
