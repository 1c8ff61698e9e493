"""Mean-teacher pieces on the GPU against vectors produced by the reference's own source
(tests/golden/ssl_golden.npz): device-side get_pseudo_labels (keep-lists bit-exact), packed ->
padded compaction, TeacherEMA (flat-buffer kernel) and the reference checkpoint layout."""
import pytest
import torch

from ssl_cases import PL_CASES, load_golden, pl_expected, pl_inputs

pytestmark = pytest.mark.gpu
G = load_golden()
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


@pytest.mark.parametrize("tag", PL_CASES)
def test_get_pseudo_labels_matches_reference(tag):
    import nesie_b200 as nb
    preds, ulb_list, ulb_flag, n_lb, n_ulb, warm = pl_inputs(G, tag, DEV)
    labels, boxes, quality = nb.get_pseudo_labels(preds, ulb_list, ulb_flag, n_lb, n_ulb,
                                                  thresh_warmup=warm)
    counts, wl, wb, wq = pl_expected(G, tag)
    assert [b.shape[0] for b in boxes] == counts
    for i in range(len(counts)):
        assert torch.equal(labels[i].float().cpu(), wl[i])
        assert torch.equal(boxes[i].cpu(), wb[i])
        assert torch.allclose(quality[i].cpu(), wq[i], rtol=1e-6, atol=1e-7)
    assert torch.equal(preds["bbox_preds"].cpu(), torch.from_numpy(G[f"{tag}_bbox_preds_after"]))


@pytest.mark.parametrize("tag", PL_CASES)
def test_packed_pseudo_labels_compact_to_the_reference_lists(tag):
    import nesie_b200 as nb
    from nesie_b200.detectors import compact_pseudo_labels
    preds, ulb_list, ulb_flag, n_lb, n_ulb, warm = pl_inputs(G, tag, DEV)
    packed = nb.get_pseudo_labels(preds, ulb_list, ulb_flag, n_lb, n_ulb, thresh_warmup=warm,
                                  as_lists=False)
    boxes, labels, valid, quality = compact_pseudo_labels(packed)
    counts, wl, wb, wq = pl_expected(G, tag)
    assert valid.sum(1).tolist() == counts
    for i, n in enumerate(counts):
        assert bool(valid[i, :n].all()) and not bool(valid[i, n:].any())
        assert torch.equal(boxes[i, :n].cpu(), wb[i])
        assert torch.equal(labels[i, :n].float().cpu(), wl[i])
        assert torch.allclose(quality[i, :n].cpu(), wq[i], rtol=1e-6, atol=1e-7)
        assert float(boxes[i, n:].abs().sum()) == 0.0


def test_teacher_ema_matches_reference_hook():
    import nesie_b200 as nb
    torch.manual_seed(int(G["seed"]) + 40)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3)).to(DEV)
    ema = nb.TeacherEMA(model, momentum=0.001, interval=1, warm_up=10)
    assert sorted(ema.ema_state_dict().keys()) == list(G["ema_buffer_names"])
    g = torch.Generator().manual_seed(int(G["seed"]) + 41)
    for it in range(4):
        with torch.no_grad():
            for p in model.parameters():
                p.add_((torch.randn(p.shape, generator=g) * 0.1).to(DEV))
        ema.after_train_iter(it)
        got = torch.cat([v.reshape(-1) for v in ema.ema_state_dict().values()]).cpu()
        want = torch.from_numpy(G["ema_after_each_step"][it])
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
    ema.swap()
    got = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()
    assert torch.allclose(got, torch.from_numpy(G["ema_swapped_params"]), rtol=1e-6, atol=1e-7)
    # reference checkpoints (`epoch_N_ema.pth`) carry the EMA copies as ema_* buffers
    state = {k: v.clone() + 1.0 for k, v in ema.ema_state_dict().items()}
    ema.load_ema_state_dict(state)
    for k, v in ema.ema_state_dict().items():
        assert torch.equal(v, state[k])
