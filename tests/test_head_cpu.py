"""The CPU oracle twin of NesieHead (oracle/nesie_head_ref.py) against golden vectors produced by the
REFERENCE's own NesieHead source (tests/golden/make_golden_head.py): targets exact, loss terms and
gradients within fp32 rounding, forward outputs (module re-created from the seed: same construction
order and state_dict names as the reference) within 1e-5."""
import pytest
import torch

from head_cases import (GRAD_KEYS, LOSS_KEYS, TARGET_NAMES, UNSUP_KEYS, load_golden, loss_inputs,
                        make_head, rel_err)
from oracle.nesie_head_ref import NesieHeadOracle

G = load_golden()


@pytest.fixture(scope="module")
def head():
    torch.manual_seed(int(G["seed"]))
    return make_head(NesieHeadOracle, 16, 8)


@pytest.mark.parametrize("tag", ["lossA", "lossB"])
def test_targets_match_reference(head, tag):
    preds, points, boxes, labels, _ = loss_inputs(G, tag, torch.device("cpu"))
    got = head.get_targets(points, boxes, labels, bbox_preds=preds)
    for name, t in zip(TARGET_NAMES, got):
        if name == "bbox_targets":
            t = torch.cat(t, dim=0)
        want = torch.from_numpy(G[f"{tag}_tgt_{name}"])
        if want.dtype in (torch.int64, torch.int32):
            assert torch.equal(t.long(), want.long()), name
        else:
            assert torch.allclose(t, want, rtol=1e-6, atol=1e-7), (name, (t - want).abs().max())


@pytest.mark.parametrize("tag", ["lossA", "lossB"])
def test_supervised_loss_and_grads_match_reference(head, tag):
    preds, points, boxes, labels, _ = loss_inputs(G, tag, torch.device("cpu"), grad=True)
    losses = head.loss(preds, points, boxes, labels)
    assert set(losses) == set(LOSS_KEYS)
    for k in LOSS_KEYS:
        assert rel_err(losses[k], G[f"{tag}_sup_{k}"]) < 2e-6, (k, losses[k].item(), G[f"{tag}_sup_{k}"])
    sum(losses.values()).backward()
    for k in GRAD_KEYS:
        g = preds[k].grad if preds[k].grad is not None else torch.zeros_like(preds[k])
        assert rel_err(g, G[f"{tag}_sup_grad_{k}"]) < 1e-5, k


@pytest.mark.parametrize("tag", ["lossA", "lossB"])
def test_unsupervised_loss_and_grads_match_reference(head, tag):
    preds, points, _, _, pl = loss_inputs(G, tag, torch.device("cpu"), grad=True)
    losses = head.unsup_loss(preds, points, pl[0], pl[1], None, pl[2])
    assert set(losses) == set(UNSUP_KEYS)
    for k in UNSUP_KEYS:
        assert rel_err(losses[k], G[f"{tag}_unsup_{k}"]) < 2e-6, (k, losses[k].item())
    sum(losses.values()).backward()
    for k in GRAD_KEYS:
        g = preds[k].grad if preds[k].grad is not None else torch.zeros_like(preds[k])
        assert rel_err(g, G[f"{tag}_unsup_grad_{k}"]) < 1e-5, k


def test_forward_matches_reference():
    B, S, C, P = (int(v) for v in G["fwd_shape"])
    torch.manual_seed(int(G["seed"]) + 10)
    head = make_head(NesieHeadOracle, C, P)
    assert sorted(head.state_dict().keys()) == list(G["fwd_keys"])
    csum = float(sum(p.detach().double().abs().sum() for p in head.parameters()))
    assert abs(csum - float(G["fwd_param_abs_sum"])) < 1e-6 * csum
    for mode in ("train", "eval"):
        head.train(mode == "train")
        feat = dict(fp_xyz=[torch.from_numpy(G["fwd_seed_points"])],
                    fp_features=[torch.from_numpy(G["fwd_seed_features"])],
                    fp_indices=[torch.from_numpy(G["fwd_seed_indices"])])
        torch.manual_seed(int(G["seed"]) + 12)
        noise = (torch.randn(B, P, 3), torch.randn(B, P, 3))
        with torch.no_grad():
            res = head(feat, "vote", "ScanNet", jitter_noise=noise)
        for k in ["vote_points", "vote_features", "aggregated_points", "aggregated_features", "obj_scores",
                  "sem_scores", "surface_pred", "surface_scale", "bbox_preds", "bbox_probs",
                  "jitter_bbox_preds", "iou_scores", "iou_scores_jitter", "side_scores",
                  "side_scores_jitter"]:
            assert rel_err(res[k], G[f"fwd_{mode}_{k}"]) < 1e-5, (mode, k)
        assert torch.equal(res["aggregated_indices"].long(),
                           torch.from_numpy(G[f"fwd_{mode}_aggregated_indices"]).long())


def test_saqe_uncertainty_weighting():
    """uncertainty='saqe' (dense_heads/saqe_head.py:590-607,631-641): the surface and IoU terms are
    weighted by exp(-sigma.detach()) with no alpha * sigma term, so side_scores receive no gradient
    from them; every other term is the Nesie head's."""
    torch.manual_seed(int(G["seed"]))
    nesie = make_head(NesieHeadOracle, 16, 8)
    saqe = make_head(NesieHeadOracle, 16, 8, uncertainty="saqe")
    preds, points, boxes, labels, _ = loss_inputs(G, "lossA", torch.device("cpu"), grad=True)
    a = nesie.loss({k: v.detach() for k, v in preds.items()}, points, boxes, labels)
    b = saqe.loss(preds, points, boxes, labels)
    for k in LOSS_KEYS:
        if k not in ("surface_loss", "iou_loss"):
            assert torch.equal(a[k], b[k]), k
    # restatement of the two SAQE lines from the same targets
    from nesie_b200.targets import pad_gt
    t = saqe.get_targets_padded(points, *pad_gt(boxes, labels, torch.device("cpu")), preds)
    C = 18
    sem = preds["sem_scores"].reshape(-1, C)
    side = preds["side_scores"].reshape(-1, 6, C)[torch.arange(sem.shape[0]), :, sem.argmax(-1)]
    sigma = 0.8 * side * side - 1.8 * side + 1
    w = t["box_loss_weights"].reshape(-1, 1).repeat(1, 6)
    from oracle.restate import bbox2surface
    L = 10.0 * ((preds["surface_pred"].reshape(-1, 6) - bbox2surface(t["bbox_targets"].reshape(-1, 7))) ** 2 * w)
    want = (torch.exp(-sigma.detach()) * L).sum()
    assert rel_err(b["surface_loss"], want.detach().numpy()) < 2e-6
    (b["surface_loss"] + b["iou_loss"]).backward()
    assert preds["side_scores"].grad is None or float(preds["side_scores"].grad.abs().sum()) == 0.0
    assert float(preds["surface_pred"].grad.abs().sum()) > 0


@pytest.mark.parametrize("per_class", [True, False])
def test_get_bboxes_matches_reference(per_class):
    """Test-time decoding (get_bboxes / multiclass_nms_single, nesie_head.py:681-788) against the
    reference's own source: objectness x IoU score, non-empty test, aligned_3d_nms over the non-empty
    boxes, score threshold, per-class expansion."""
    from head_cases import check_decoded, decode_inputs
    torch.manual_seed(0)
    head = make_head(NesieHeadOracle, 16, 8, test_cfg=dict(nms_thr=0.25, score_thr=0.05, per_class_proposal=per_class))
    head = head.to(torch.device("cpu"))
    points, preds = decode_inputs(G, torch.device("cpu"))
    check_decoded(G, per_class, head.get_bboxes(points, preds))
