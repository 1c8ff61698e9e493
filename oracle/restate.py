"""numpy / torch-CPU restatements of the python-level pieces of the hot path.

TEST INFRASTRUCTURE ONLY.  Each function names the reference lines it follows; arithmetic is
written so that every python operator of the reference is one rounding here too.  These are
checked against vectors produced by the reference's own source (tests/golden/make_golden_py.py
execs the reference's function bodies from /root/reference) before they are trusted.
"""
import numpy as np
import torch


def aligned_3d_nms(boxes, scores, classes, thresh):
    """core/post_processing/box3d_nms.py:129-176 in numpy fp32.  Stable ascending sort
    (the reference's argsort is unstable: only distinct scores are comparable)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    classes = np.asarray(classes)
    lo, hi = boxes[:, :3], boxes[:, 3:6]
    ext = hi - lo
    area = ext[:, 0] * ext[:, 1] * ext[:, 2]
    order = np.argsort(scores, kind="stable")
    thresh = np.float32(thresh)
    pick = []
    with np.errstate(divide="ignore", invalid="ignore"):
        while order.size:
            i, rest = order[-1], order[:-1]
            pick.append(int(i))
            a = np.maximum(lo[i], lo[rest])
            b = np.minimum(hi[i], hi[rest])
            d = np.maximum(np.float32(0), b - a)
            inter = d[:, 0] * d[:, 1] * d[:, 2]
            iou = inter / (area[i] + area[rest] - inter)
            iou = iou * (classes[i] == classes[rest]).astype(np.float32)
            order = rest[iou <= thresh]
    return np.asarray(pick, dtype=np.int64)


def lhs_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """models/detectors/votenet_nesie.py:733-779 in numpy fp64 (stable sort)."""
    boxes = np.asarray(boxes, dtype=np.float64)
    lo, hi, score, cls = boxes[:, :3], boxes[:, 3:6], boxes[:, 6], boxes[:, 7]
    ext = hi - lo
    area = ext[:, 0] * ext[:, 1] * ext[:, 2] + 1e-8
    order = np.argsort(score, kind="stable")
    pick = []
    with np.errstate(divide="ignore", invalid="ignore"):
        while order.size:
            i, rest = order[-1], order[:-1]
            pick.append(int(i))
            a = np.maximum(lo[i], lo[rest])
            b = np.minimum(hi[i], hi[rest])
            d = np.maximum(0, b - a)
            inter = d[:, 0] * d[:, 1] * d[:, 2]
            o = inter / area[rest] if old_type else inter / (area[i] + area[rest] - inter)
            o = o * (cls[i] == cls[rest])
            hit = np.where(o > overlap_threshold)[0]
            for t in range(len(hit) // 2):  # top half of the suppressed set is kept too
                pick.append(int(rest[hit[len(hit) - 1 - t]]))
            order = np.delete(rest, hit)
    return pick


def bbox2surface(bbox):
    """models/losses/surface_loss.py:90-100"""
    c, s = bbox[..., :3], bbox[..., 3:6]
    return torch.cat([c - 0.5 * s, c + 0.5 * s], dim=-1)


def side_uncertainty_loss(surface_pred, box_targets, side_scores, sem_scores, weight,
                          loss_weight=10.0, alpha=1.0):
    """models/dense_heads/nesie_head.py:332-349 with SurfaceLoss(MSE, reduction 'none',
    loss_weight) of models/losses/surface_loss.py:57-61 and mmdet's weighted mse_loss.
    Differentiable torch code (autograd supplies the reference gradients)."""
    target = bbox2surface(box_targets)
    loss = torch.nn.functional.mse_loss(surface_pred, target, reduction="none")
    loss = loss_weight * (loss * weight)
    indx = sem_scores.max(dim=-1)[1].reshape(-1)
    rows = indx.shape[0]
    side = side_scores[torch.arange(rows, device=indx.device), :, indx].reshape(-1, 6)
    sigma = 0.8 * side * side - 1.8 * side + torch.ones_like(side)
    out = torch.exp(-sigma) * loss + alpha * sigma * weight
    return out.sum(), sigma


def ema_update(ema, param, momentum, warm_up, curr_step):
    """core/utils/simi_teacher_hook.py:54-64 (in place on `ema`)."""
    m = min(momentum, (1 + curr_step) / (warm_up + curr_step))
    ema.mul_(1 - m).add_(param, alpha=m)
    return ema


# ---- pseudo-label filter, loop form ----------------------------------------------------------

def _get_3d_box_minmax(size, center_cam):
    """votenet_nesie.py:790-813 with heading 0, reduced to the corner min/max the caller takes
    (:245-250).  fp32 corners: fl32(centre +- extent/2); extents are (l, h, w) on (X, Y, Z)."""
    l, w, h = [np.float32(v) for v in size]
    half = np.array([l / 2, h / 2, w / 2], dtype=np.float32)
    c = np.asarray(center_cam, dtype=np.float32)
    lo, hi = (c - half).astype(np.float32), (c + half).astype(np.float32)
    return np.minimum(lo, hi), np.maximum(lo, hi)


def get_pseudo_labels(preds, ulb_list, ulb_flag, n_lb, n_ulb, num_classes=18, thresh_warmup=True,
                      use_cbl=True, quality_poly=(5 / 3, 8 / 3), nms_iou=0.25):
    """models/detectors/votenet_nesie.py:129-299 on CPU tensors, proposal-by-proposal loops kept
    where the reference loops (including its two indexing quirks, see
    nesie_b200/pseudo_label.py).  Stable sorts.  Returns the three per-scene lists."""
    preds = {k: v.detach().cpu().clone() for k, v in preds.items()}
    counter = ulb_list.sum(dim=0)
    srt, order = torch.sort(counter, descending=True)
    acc = torch.zeros(num_classes)
    ulb_count = 10 * ulb_flag.sum() * n_lb / n_ulb
    for i in order.tolist():
        denom = max(max(srt), ulb_count) if thresh_warmup else max(srt)
        acc[i] = srt[i] / denom
        acc[i] = acc[i] / (2.0 - acc[i])

    bbox = preds["bbox_preds"]
    bbox[:, :, 2] = bbox[:, :, 2] - bbox[:, :, 5] * 0.5
    center, size, heading = bbox[:, :, :3], bbox[:, :, 3:6], bbox[:, :, 6:7]
    B, P = center.shape[:2]
    max_cls, argmax_cls = torch.max(preds["sem_scores"], dim=2)
    flat = argmax_cls.reshape(-1)
    if use_cbl:
        threshold = torch.tensor([float(acc[flat[int(i)]]) for i in flat]).reshape(B, P)
        cls_thr = 0.7 + 0.3 * threshold
        cls_thr[cls_thr > 0.95] = 0.95
    else:
        cls_thr = 0.9
    cls_mask = max_cls > cls_thr
    obj = torch.softmax(preds["obj_scores"], dim=2)
    pos_obj, neg_obj = obj[:, :, 1], obj[:, :, 0]
    obj_mask, neg_mask = pos_obj > 0.9, neg_obj > 0.9
    iou_logits = preds["iou_scores"].reshape(-1, num_classes)
    iou_pred = torch.stack([iou_logits[i, flat[i]] for i in range(B * P)]).reshape(B, -1)
    if use_cbl:
        iou_thr = 0.25 + threshold * 0.5
        iou_thr[iou_thr > 0.35] = 0.35
    else:
        iou_thr = 0.25
    final_mask = cls_mask & obj_mask & (iou_pred > iou_thr)
    side = preds["side_scores"].reshape(-1, 6, num_classes)
    side_sel = torch.stack([side[i, :, flat[i]] for i in range(B * P)]).reshape(B, -1, 6)
    qa, qb = quality_poly
    quality = qa * side_sel * side_sel - qb * side_sel + torch.ones_like(side_sel)

    K = 64
    inds = torch.argsort(pos_obj * iou_pred * final_mask, dim=1, descending=True, stable=True)
    inds = inds[:, :K]
    K = inds.shape[1]
    fm_sorted = torch.gather(final_mask, 1, inds)
    neg_sorted = torch.gather(neg_mask, 1, inds)
    for b in range(B):
        rows = np.zeros((K, 8))
        for j in range(K):
            p = int(inds[b, j])
            c = center[b, p].numpy()
            cam = np.array([c[0], -c[2], c[1]], dtype=np.float32)  # flip_axis_to_camera :781-788
            lo, hi = _get_3d_box_minmax(size[b, p].numpy(), cam)
            rows[j, 0:3], rows[j, 3:6] = lo, hi
            rows[j, 6] = pos_obj[b, p].numpy() * iou_pred[b, p].numpy()
            rows[j, 7] = int(argmax_cls[b, p])
        pick = lhs_3d_faster_samecls(rows, nms_iou, False)
        keep = np.zeros(K, dtype=bool)
        keep[pick] = True
        fm_sorted[b] &= torch.from_numpy(keep)

    labels, boxes, quals = [], [], []
    for b in range(B):
        sel = [j for j in range(K) if fm_sorted[b, j]]
        if sel:
            p = inds[b, sel]
            labels.append(argmax_cls[b, p])
            boxes.append(torch.cat([center[b, p], size[b, p], heading[b, p]], dim=-1))
            quals.append(quality[b, p])
        else:
            labels.append(torch.ones((0,)))
            boxes.append(torch.ones((0, 7)))
            quals.append(torch.ones((0, 6)))
    return labels, boxes, quals
