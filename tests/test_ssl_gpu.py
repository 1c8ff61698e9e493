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


def test_scene_batcher_and_views_on_the_device(tmp_path):
    """SURVEY 8f-4 on the GPU: `.bin` scenes -> pinned host -> device sampling -> the two augmented
    views of a mean-teacher batch; a box centre moves like the point at its centre."""
    import numpy as np
    from nesie_b200 import data
    from nesie_b200.detectors import transform_boxes
    files = []
    rng = np.random.default_rng(0)
    for i in range(4):
        f = str(tmp_path / f"s{i}.bin")
        data.save_points(f, rng.normal(size=(60000 if i % 2 else 30000, 6)).astype(np.float32))
        files.append(f)
    pts, choices = next(iter(data.SceneBatcher(files, num_points=40000, batch_size=4, device=DEV)))
    assert pts.shape == (4, 40000, 4) and pts.is_cuda
    assert choices[1].unique().numel() == 40000 and choices[0].unique().numel() < 40000
    ref = torch.from_numpy(data.load_points(files[1])).to(DEV)
    assert torch.equal(pts[1], ref[choices[1]])
    boxes = torch.rand(4, 6, 7, device=DEV)
    v = data.mean_teacher_views(pts, boxes, torch.Generator().manual_seed(2))
    assert torch.allclose(transform_boxes(boxes, v["aug_s"])[..., :3],
                          v["aug_s"].apply_points(boxes[..., :3]), atol=1e-6)
    assert torch.equal(v["points_s"][..., 3], pts[..., 3])
