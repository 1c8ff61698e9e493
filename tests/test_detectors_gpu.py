"""Whole-step parity on the GPU: nesie_b200.detectors.VoteNet / VoteNetNesie (every kernel of this
repo in the loop) against their CPU twins (oracle/detectors_ref.py: C restatement of the reference
kernels, torch-CPU MLPs, loop-form pseudo-label filter), same weights, same jitter noise, same data.
Loss terms within 1e-5 of the loss scale (the north star's fp32 bar), pseudo-label keep-lists exact,
weights after optimizer + EMA step within 1e-5."""
import copy
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_distributed_cpu import tiny_votenet  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _liven(model):
    """A randomly initialised tiny model has no proposal near a GT centre and none that passes the
    pseudo-label thresholds: widen the assignment radius and bias objectness / one class so that
    every loss term and the filter are exercised."""
    model.train_cfg.update(pos_distance_thr=1.2, neg_distance_thr=2.0)
    bias = model.bbox_head.conv_pred.conv_cls.bias.data
    bias[1] += 4.0
    bias[2 + 3] += 2.5


def _noise(B, P, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, P, 3, generator=g), torch.randn(B, P, 3, generator=g)


def test_votenet_pretrain_step_matches_cpu_twin():
    from nesie_b200 import targets as T
    from nesie_b200.detectors import VoteNet
    from nesie_b200.synthetic import make_batch
    from oracle.detectors_ref import VoteNetRef
    _setup()
    torch.manual_seed(3)
    ref = tiny_votenet(VoteNetRef)
    gpu = tiny_votenet(VoteNet)
    _liven(ref)
    _liven(gpu)
    gpu.load_state_dict(ref.state_dict())
    gpu = gpu.to(DEV)
    pts, gb, gl = make_batch(3, 2048, seed0=40, origin="bottom")
    gb[1], gl[1] = gb[1][:0], gl[1][:0]                       # a scene without GT boxes
    n1, n2 = _noise(3, 16, 5)
    want = ref.bbox_head.loss(ref.predict(pts, jitter_noise=(n1, n2)), pts, gb, gl)
    preds = gpu.predict(pts.to(DEV), jitter_noise=(n1.to(DEV), n2.to(DEV)))
    got = gpu.bbox_head.loss_padded(preds, pts.to(DEV), *T.pad_gt(gb, gl, DEV, pad_to=16))
    scale = max(float(v.abs()) for v in want.values())
    for k in want:
        assert abs(float(got[k]) - float(want[k])) < 1e-5 * max(scale, 1.0), (k, float(got[k]), float(want[k]))
    sum(want.values()).backward()
    sum(got.values()).backward()
    gw = torch.cat([p.grad.reshape(-1) for p in ref.parameters() if p.grad is not None])
    gg = torch.cat([p.grad.reshape(-1) for p, q in zip(gpu.parameters(), ref.parameters())
                    if q.grad is not None]).cpu()
    cos = torch.nn.functional.cosine_similarity(gw, gg, dim=0)
    assert cos > 0.9999, float(cos)


@pytest.mark.parametrize("teacher_mode", ["overlap", "swap"])
def test_mean_teacher_step_matches_cpu_twin(teacher_mode):
    from nesie_b200 import targets as T
    from nesie_b200.detectors import BoxAug, VoteNetNesie, transform_boxes
    from nesie_b200.synthetic import make_batch
    from oracle.detectors_ref import VoteNetNesieRef
    _setup()
    torch.manual_seed(4)
    kw = dict(n_lb=6, n_ulb=20)
    ref = tiny_votenet(lambda **k: VoteNetNesieRef(**k, **kw))
    gpu = tiny_votenet(lambda **k: VoteNetNesie(**k, teacher_mode=teacher_mode, **kw))
    for m in (ref, gpu):
        _liven(m)
        m.train_cfg.update(use_cbl=False)
    gpu.load_state_dict(ref.state_dict())
    gpu = gpu.to(DEV)
    ref.init_teacher()
    gpu.init_teacher()
    S, nl = 4, 2
    pts, gb, gl = make_batch(S, 2048, seed0=60, origin="bottom")
    g = torch.Generator().manual_seed(9)
    aug_s, aug_t = BoxAug.random(S, torch.device("cpu"), g), BoxAug.random(S, torch.device("cpu"), g)
    b, l, v = T.pad_gt(gb[:nl], gl[:nl], torch.device("cpu"), pad_to=16)
    b = transform_boxes(b, aug_s.index(torch.arange(nl))) * v.unsqueeze(-1)
    ps, pt = aug_s.apply_points(pts), aug_t.apply_points(pts)
    sup, unsup = torch.arange(nl), torch.arange(nl, S)
    pos = torch.tensor([3, 11])
    n1, n2 = _noise(S, 16, 6)

    def run(m, dev):
        d = lambda t: t.to(dev)   # noqa: E731
        mv = lambda a: BoxAug(d(a.hf), d(a.vf), d(a.rot), d(a.scale), d(a.trans))   # noqa: E731
        jit = dict(jitter_noise=(d(n1), d(n2)))
        return m.forward_train_padded(d(ps), d(pt), d(b), d(l), d(v), d(sup), d(unsup), d(pos),
                                      mv(aug_t), mv(aug_s), student_kw=jit, teacher_kw=jit)
    want = run(ref, torch.device("cpu"))
    got = run(gpu, DEV)
    assert set(want) == set(got) and len(want) == 12
    scale = max(float(x.abs()) for x in want.values())
    for k in want:
        assert abs(float(got[k]) - float(want[k])) < 1e-5 * max(scale, 1.0), (k, float(got[k]), float(want[k]))
    assert torch.equal(gpu.ulb_list.cpu(), ref.ulb_list) and torch.equal(gpu.ulb_flag.cpu(), ref.ulb_flag)
    # BatchNorm running statistics: written by the student AND the teacher pass (in that order)
    for (n1, b1), (n2, b2) in zip(gpu.named_buffers(), ref.named_buffers()):
        assert n1 == n2
        if n1.endswith("num_batches_tracked"):
            assert int(b1) == int(b2) == 2, n1
        elif "running_" in n1:
            assert float((b1.cpu() - b2).abs().max()) < 1e-5 * max(float(b2.abs().max()), 1.0), n1
    # optimizer + EMA: weights and teacher copies stay together
    for m, losses in ((ref, want), (gpu, got)):
        opt = torch.optim.AdamW(m.parameters(), lr=0.008, weight_decay=0.01)
        sum(losses.values()).backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 10.0)
        opt.step()
        m.after_train_iter(0)
    w_ref = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    w_gpu = torch.cat([p.detach().reshape(-1) for p in gpu.parameters()]).cpu()
    # AdamW's first step moves every weight by ~lr * g / (|g| + eps): where a gradient is within
    # rounding of zero the two sides may step in different directions, everywhere else they agree
    close = (w_ref - w_gpu).abs() < 1e-5 * w_ref.abs().max()
    assert close.float().mean() > 0.99, float(close.float().mean())
    g_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    g_gpu = torch.cat([p.grad.reshape(-1) for p in gpu.parameters()]).cpu()
    assert torch.nn.functional.cosine_similarity(g_ref, g_gpu, dim=0) > 0.9999
    e_ref = torch.cat([e.reshape(-1) for e in ref.teacher.ema])
    e_gpu = torch.cat([v.reshape(-1) for v in gpu.teacher.ema_state_dict().values()]).cpu()
    assert ((e_ref - e_gpu).abs() < 1e-5 * e_ref.abs().max()).float().mean() > 0.99
