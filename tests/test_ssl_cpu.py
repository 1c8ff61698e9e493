"""CPU side of the mean-teacher pieces against vectors produced by the reference's own source
(tests/golden/make_golden_ssl.py): the oracle restatements of get_pseudo_labels and the EMA hook,
and the torch glue of nesie_b200.detectors (box transforms, class-count table, row selection), which
runs on whatever device its tensors live on."""
import numpy as np
import pytest
import torch

from oracle import restate
from ssl_cases import PL_CASES, aug_from_golden, load_golden, padded_boxes, pl_expected, pl_inputs

G = load_golden()
CPU = torch.device("cpu")


@pytest.mark.parametrize("tag", PL_CASES)
def test_oracle_get_pseudo_labels_matches_reference(tag):
    preds, ulb_list, ulb_flag, n_lb, n_ulb, warm = pl_inputs(G, tag, CPU)
    labels, boxes, quality = restate.get_pseudo_labels(preds, ulb_list, ulb_flag, n_lb, n_ulb,
                                                       thresh_warmup=warm)
    counts, wl, wb, wq = pl_expected(G, tag)
    assert [b.shape[0] for b in boxes] == counts
    for i in range(len(counts)):
        assert torch.equal(labels[i].float(), wl[i])
        assert torch.equal(boxes[i], wb[i])
        assert torch.equal(quality[i], wq[i])


def test_oracle_ema_matches_reference_hook():
    torch.manual_seed(int(G["seed"]) + 40)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    ema = [p.detach().clone() for p in model.parameters()]
    g = torch.Generator().manual_seed(int(G["seed"]) + 41)
    for it in range(4):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
        for e, p in zip(ema, model.parameters()):
            restate.ema_update(e, p.detach(), 0.001, 10, it)
        flat = torch.cat([e.reshape(-1) for e in ema])
        assert torch.equal(flat, torch.from_numpy(G["ema_after_each_step"][it]))


def test_box_transform_matches_reference():
    from nesie_b200.detectors import transformation_bbox_preds
    boxes, counts = padded_boxes(G, CPU)
    out = transformation_bbox_preds(boxes, aug_from_golden(G, "t", CPU), aug_from_golden(G, "s", CPU))
    got = torch.cat([out[i, :n] for i, n in enumerate(counts)])
    want = torch.from_numpy(G["tf_result"])
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6), (got - want).abs().max()


def test_points_and_boxes_transform_consistently():
    """A box centre transformed as a box lands where the same point lands as a point."""
    from nesie_b200.detectors import BoxAug, transform_boxes, untransform_boxes
    g = torch.Generator().manual_seed(3)
    aug = BoxAug.random(5, CPU, g, scale_range=(0.9, 1.1), trans_std=0.1)
    boxes = torch.cat([torch.randn(5, 7, 3, generator=g), torch.rand(5, 7, 3, generator=g) + 0.2,
                       torch.zeros(5, 7, 1)], -1)
    moved = transform_boxes(boxes, aug)
    assert torch.allclose(moved[..., :3], aug.apply_points(boxes[..., :3]), atol=1e-6)
    back = untransform_boxes(moved, aug)
    assert torch.allclose(back[..., :6], boxes[..., :6], atol=1e-5)


@pytest.mark.parametrize("tag", PL_CASES)
def test_ulb_update_matches_reference(tag):
    from nesie_b200.detectors import ulb_update
    _, ulb_list, ulb_flag, _, _, _ = pl_inputs(G, tag, CPU)
    counts, labels, _, _ = pl_expected(G, tag)
    Gmax = 64
    lab = torch.zeros(len(counts), Gmax, dtype=torch.long)
    valid = torch.zeros(len(counts), Gmax, dtype=torch.bool)
    for i, n in enumerate(counts):
        lab[i, :n] = labels[i].long()
        valid[i, :n] = True
    ulb_update(ulb_list, ulb_flag, torch.from_numpy(G[f"{tag}_ulb_pos"]), lab, valid)
    assert torch.equal(ulb_list, torch.from_numpy(G[f"{tag}_ulb_list_after"]))
    assert torch.equal(ulb_flag, torch.from_numpy(G[f"{tag}_ulb_flag_after"]))


def test_choose_items_matches_reference():
    from nesie_b200.detectors import choose_items
    use = torch.from_numpy(G["choose_use_label"])
    preds = dict(a=torch.from_numpy(G["choose_a"]), b=torch.from_numpy(G["choose_b"]))
    sup = choose_items(preds, torch.nonzero(use).squeeze(1))
    unsup = choose_items(preds, torch.nonzero(~use).squeeze(1))
    assert torch.equal(sup["a"], torch.from_numpy(G["choose_sup_a"]))
    assert torch.equal(unsup["a"], torch.from_numpy(G["choose_unsup_a"]))
