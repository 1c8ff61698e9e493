"""Mean-teacher filter ops at the BASELINE config-4 shape: 16 scenes x 256
proposals.  This repo's batched kernels vs the reference-style formulation (per-scene python loop of
small torch ops for aligned_3d_nms, host numpy for the lenient NMS, per-tensor EMA loop)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesie_b200 as nb  # noqa: E402
from oracle import restate  # noqa: E402


def gpu_time(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def torch_loop_nms(boxes, scores, classes, thresh):
    """aligned_3d_nms the way the reference runs it: a python while-loop of tiny torch ops on the
    GPU (restated; semantics of core/post_processing/box3d_nms.py:129-176)."""
    lo, hi = boxes[:, :3], boxes[:, 3:]
    area = (hi - lo).prod(-1)
    order = torch.argsort(scores)
    pick = []
    while order.shape[0] != 0:
        i, rest = order[-1], order[:-1]
        pick.append(i)
        d = (torch.min(hi[i], hi[rest]) - torch.max(lo[i], lo[rest])).clamp(min=0)
        inter = d[:, 0] * d[:, 1] * d[:, 2]
        iou = inter / (area[i] + area[rest] - inter) * (classes[i] == classes[rest]).float()
        order = rest[torch.nonzero(iou <= thresh, as_tuple=False).flatten()]
    return torch.stack(pick) if pick else order


def main():
    S, P, C = 16, 256, 18
    rng = np.random.default_rng(0)
    anchors = rng.uniform(-4, 4, (S, 20, 3))
    c = anchors[np.arange(S)[:, None], rng.integers(0, 20, (S, P))] + rng.normal(0, 0.2, (S, P, 3))
    sz = rng.uniform(0.2, 1.5, (S, P, 3))
    boxes = torch.tensor(np.concatenate([c - sz / 2, c + sz / 2], -1), dtype=torch.float32).cuda()
    scores = torch.rand(S, P).cuda()
    classes = torch.randint(0, C, (S, P)).cuda()
    out = []
    t_mine = gpu_time(lambda: nb.aligned_3d_nms_batched(boxes, scores, classes, 0.25))
    t_ref = gpu_time(lambda: [torch_loop_nms(boxes[s], scores[s], classes[s], 0.25) for s in range(S)], 2)
    out.append({"op": "aligned_3d_nms 16x256", "ms": round(t_mine, 4), "reference_style_ms": round(t_ref, 2),
                "speedup": round(t_ref / t_mine, 1)})
    rows = torch.zeros(S, 64, 8, dtype=torch.float64).cuda()
    rows[..., :6] = boxes[:, :64].double()
    rows[..., 6] = scores[:, :64].double()
    rows[..., 7] = classes[:, :64].double()
    t_mine = gpu_time(lambda: nb.lhs_3d_faster_samecls_batched(rows, 0.25))

    def host_lhs():
        r = rows.cpu().numpy()
        return [restate.lhs_3d_faster_samecls(r[s], 0.25) for s in range(S)]
    t_ref = gpu_time(host_lhs, 3)
    out.append({"op": "lhs_nms 16x64 (fp64)", "ms": round(t_mine, 4), "reference_style_ms": round(t_ref, 2),
                "speedup": round(t_ref / t_mine, 1)})
    preds = dict(bbox_preds=torch.cat([torch.tensor(c, dtype=torch.float32), torch.tensor(sz, dtype=torch.float32),
                                       torch.zeros(S, P, 1)], -1).cuda(),
                 sem_scores=(torch.rand(S, P, C) ** 0.3).cuda(), obj_scores=(torch.randn(S, P, 2) * 5).cuda(),
                 iou_scores=torch.rand(S, P, C).cuda(), side_scores=torch.rand(S, P, 6, C).cuda(),
                 vote_points=torch.rand(S, P, 3).cuda())
    ulb_list, ulb_flag = torch.randint(0, 6, (100, C)).float().cuda(), torch.ones(100).cuda()
    t_mine = gpu_time(lambda: nb.get_pseudo_labels({k: v.clone() for k, v in preds.items()}, ulb_list, ulb_flag,
                                                   12, 100), 5)
    cpu_preds = {k: v.cpu() for k, v in preds.items()}
    t0 = time.perf_counter()
    restate.get_pseudo_labels(cpu_preds, ulb_list.cpu(), ulb_flag.cpu(), 12, 100)
    t_ref = (time.perf_counter() - t0) * 1e3
    out.append({"op": "get_pseudo_labels 16x256", "ms": round(t_mine, 3), "reference_style_ms": round(t_ref, 1),
                "speedup": round(t_ref / t_mine, 1), "note": "reference style = python loops on the host"})
    rowsN = S * P
    args = (torch.randn(rowsN, 6).cuda().requires_grad_(True), torch.rand(rowsN, 7).cuda(),
            torch.rand(rowsN, 6, C).cuda().requires_grad_(True), torch.randn(rowsN, C).cuda(),
            torch.rand(rowsN, 6).cuda())

    def mine_loss():
        l, s = nb.side_uncertainty_loss(*args)
        l.backward()

    def ref_loss():
        l, s = restate.side_uncertainty_loss(*args)
        l.backward()
    t_mine, t_ref = gpu_time(mine_loss), gpu_time(ref_loss)
    out.append({"op": "side_uncertainty_loss fwd+bwd 4096 rows", "ms": round(t_mine, 4),
                "reference_style_ms": round(t_ref, 3), "speedup": round(t_ref / t_mine, 1)})
    model = nb.PointNet2SASSG(in_channels=4).cuda()
    ema = nb.TeacherEMA(model)
    params = list(model.parameters())
    bufs = [p.detach().clone() for p in params]
    t_mine = gpu_time(lambda: ema.after_train_iter(100))

    def ref_ema():
        for b, p in zip(bufs, params):
            b.mul_(0.999).add_(p.data, alpha=0.001)
    t_ref = gpu_time(ref_ema)
    out.append({"op": f"teacher EMA ({sum(p.numel() for p in params)} params, {len(params)} tensors)",
                "ms": round(t_mine, 4), "reference_style_ms": round(t_ref, 3), "speedup": round(t_ref / t_mine, 1)})
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
