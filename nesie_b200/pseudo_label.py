"""Mean-teacher pseudo-label filter on the device.

Mirror of VoteNetNesie.get_pseudo_labels / lhs_3d_faster_samecls
(reference: mmdet3d/models/detectors/votenet_nesie.py:129-299, 733-821; the SAQE twin differs only
in the quality polynomial, votenet_saqe.py:201).  The reference pulls the teacher's predictions to
the host (>= 7 .cpu() copies), builds box corners in a B x 64 python loop and runs a numpy NMS per
scene; here every step stays on the GPU: vectorised gathers replace the python list
comprehensions and the lenient float64 NMS runs for all scenes in one launch
(nesie_lhs_nms_batched).

Two quirks of the reference are reproduced on purpose, because the results must be identical:
  * votenet_nesie.py:143-149 fills `classwise_acc[i] = sorted[i] / ...` for i in `indices`, i.e.
    class i receives the i-th LARGEST pseudo-label count, not its own count;
  * votenet_nesie.py:161 builds `threshold[p] = classwise_acc[argmax_flat[argmax_flat[p]]]` (it
    iterates over the class ids and uses them as positions into the flattened argmax array).
The top-64 selection uses a STABLE descending sort (the reference's torch.argsort is unstable, so
its filler entries among the masked-out proposals are implementation-defined).
"""
import torch

from . import _lib

MAX_NUM_OBJ = 64


def lhs_3d_faster_samecls_batched(boxes, overlap_threshold, old_type=False, counts=None):
    """boxes (S, n, 8) float64 rows [x1,y1,z1,x2,y2,z2,score,cls] -> (pick (S, 2n) int32 padded
    with -1, pick_cnt (S,) int32), pick order as votenet_nesie.py:733-779."""
    _lib.need_cuda(boxes)
    S, n, _ = boxes.shape
    boxes = boxes.contiguous().double()
    if counts is not None:
        counts = counts.contiguous().to(torch.int32)
    pick = torch.full((S, 2 * n), -1, dtype=torch.int32, device=boxes.device)
    pick_cnt = torch.zeros((S,), dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        _lib.call("nesie_lhs_nms_batched", S, n, _lib.ptr(boxes), _lib.ptr(counts),
                  float(overlap_threshold), int(bool(old_type)), _lib.ptr(pick),
                  _lib.ptr(pick_cnt), _lib.stream())
    return pick, pick_cnt


def lhs_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """Single-scene form with the reference's signature: (n, 8) -> python list of picks."""
    boxes = torch.as_tensor(boxes)
    if boxes.shape[0] == 0:
        return []
    pick, cnt = lhs_3d_faster_samecls_batched(boxes[None].cuda(), overlap_threshold, old_type)
    return pick[0, :int(cnt.item())].tolist()


def classwise_acc_from_counts(ulb_list, ulb_flag, n_lb, n_ulb, thresh_warmup=True):
    """votenet_nesie.py:133-149 (including the sorted-position quirk described above)."""
    pseudo_counter = ulb_list.sum(dim=0)
    sorted_cnt, _ = torch.sort(pseudo_counter, descending=True)
    top = sorted_cnt.max()
    if thresh_warmup:
        ulb_count = 10 * ulb_flag.sum() * n_lb / n_ulb
        top = torch.maximum(top, ulb_count.to(top.dtype))
    acc = sorted_cnt / top
    return acc / (2.0 - acc)


def corners_minmax_camera(center, size):
    """Axis-aligned (min, max) of the 8 corners get_3d_box builds with heading 0 in the
    'upright camera' frame (votenet_nesie.py:781-813: X = x, Y = -z, Z = y; extents l, h, w).
    center/size (..., 3) fp32 -> (..., 6) fp32; every corner is fl32(centre +- extent/2)."""
    c = torch.stack([center[..., 0], -center[..., 2], center[..., 1]], dim=-1)
    half = torch.stack([size[..., 0], size[..., 2], size[..., 1]], dim=-1) / 2
    lo, hi = c - half, c + half
    return torch.cat([torch.minimum(lo, hi), torch.maximum(lo, hi)], dim=-1)


def get_pseudo_labels(unsup_bbox_preds, ulb_list, ulb_flag, n_lb, n_ulb, num_classes=18,
                      thresh_warmup=True, use_cbl=True, dataset_name="ScanNet",
                      quality_poly=(5 / 3, 8 / 3), nms_iou=0.25, as_lists=True):
    """Device-side get_pseudo_labels.

    unsup_bbox_preds: dict with bbox_preds (B,P,7), sem_scores (B,P,C), obj_scores (B,P,2),
    iou_scores (B,P,C), side_scores (B,P,6,C), vote_points (B,P,3).  ulb_list (n_ulb, C) and
    ulb_flag (n_ulb,) are the class-count tables of the reference's runner hook.
    quality_poly = (a, b) of quality = a s^2 - b s + 1 (Nesie 5/3, 8/3; SAQE 0.8, 1.8).

    Returns (pseudo_label, pseudo_boxes, pseudo_quality_score): per-scene lists when as_lists
    (one host sync for the variable lengths) or the packed (B,64,...) tensors + label_mask.
    Like the reference, bbox_preds[..., 2] is shifted down by half the box height in place."""
    preds = unsup_bbox_preds
    dev = preds['sem_scores'].device
    classwise_acc = classwise_acc_from_counts(ulb_list.to(dev), ulb_flag.to(dev), n_lb, n_ulb,
                                              thresh_warmup).float()
    bbox = preds['bbox_preds']
    bbox[:, :, 2] = bbox[:, :, 2] - bbox[:, :, 5] * 0.5
    pred_center, pred_size, pred_heading = bbox[:, :, :3], bbox[:, :, 3:6], bbox[:, :, 6:7]
    B, P = pred_center.shape[:2]

    max_cls, argmax_cls = torch.max(preds['sem_scores'], dim=2)
    flat = argmax_cls.reshape(-1)
    if use_cbl:
        threshold = classwise_acc[flat[flat]].reshape(B, P)
        cls_threshold = (0.7 + 0.3 * threshold).clamp(max=0.95)
    else:
        threshold = None
        cls_threshold = 0.9
    cls_mask = max_cls > cls_threshold

    pred_objectness = torch.softmax(preds['obj_scores'], dim=2)
    pos_obj, neg_obj = pred_objectness[:, :, 1], pred_objectness[:, :, 0]
    obj_threshold = 0.9
    objectness_mask = pos_obj > obj_threshold
    neg_objectness_mask = neg_obj > obj_threshold

    iou_pred = torch.gather(preds['iou_scores'], 2, argmax_cls.unsqueeze(-1)).squeeze(-1)
    if use_cbl:
        iou_threshold = (0.25 + threshold * 0.5).clamp(max=0.35)
    else:
        iou_threshold = 0.25
    final_mask = cls_mask & objectness_mask & (iou_pred > iou_threshold)

    side = preds['side_scores'].detach()
    sel = argmax_cls[:, :, None, None].expand(B, P, 6, 1)
    side_scores = torch.gather(side, 3, sel).squeeze(-1)
    qa, qb = quality_poly
    quality_score = qa * side_scores * side_scores - qb * side_scores + torch.ones_like(side_scores)

    inds = torch.argsort(pos_obj * iou_pred * final_mask, dim=1, descending=True, stable=True)
    inds = inds[:, :MAX_NUM_OBJ]
    K = inds.shape[1]
    i3 = inds.unsqueeze(-1).expand(-1, -1, 3)
    final_mask_sorted = torch.gather(final_mask, 1, inds)
    neg_objectness_mask = torch.gather(neg_objectness_mask, 1, inds)

    center_ = torch.gather(pred_center, 1, i3).detach()
    size_ = torch.gather(pred_size, 1, i3).detach()
    boxes = torch.empty((B, K, 8), dtype=torch.float64, device=dev)
    boxes[:, :, :6] = corners_minmax_camera(center_, size_).double()
    boxes[:, :, 6] = (torch.gather(pos_obj, 1, inds) * torch.gather(iou_pred, 1, inds)).detach().double()
    boxes[:, :, 7] = torch.gather(argmax_cls, 1, inds).double()
    pick, pick_cnt = lhs_3d_faster_samecls_batched(boxes, nms_iou, False)
    valid = torch.arange(pick.shape[1], device=dev)[None, :] < pick_cnt[:, None]
    picked = torch.zeros((B, K + 1), dtype=torch.bool, device=dev)
    picked.scatter_(1, torch.where(valid, pick.long(), torch.full_like(pick, K).long()), True)
    final_mask_sorted = final_mask_sorted & picked[:, :K]

    label_mask = final_mask_sorted.long()
    heading_label = torch.gather(pred_heading, 1, inds.unsqueeze(-1))
    size_label = torch.gather(pred_size, 1, i3)
    quality_score = torch.gather(quality_score, 1, inds.unsqueeze(-1).expand(-1, -1, 6))
    sem_cls_label = torch.gather(argmax_cls, 1, inds)
    center_label = torch.gather(pred_center, 1, i3).clone()
    center_label[~final_mask_sorted] = -1000
    packed = dict(label_mask=label_mask, center_label=center_label, size_label=size_label,
                  heading_label=heading_label, sem_cls_label=sem_cls_label,
                  quality_score=quality_score, neg_objectness_mask=neg_objectness_mask,
                  inds=inds)
    if not as_lists:
        return packed
    box_label = torch.cat([center_label, size_label, heading_label], dim=-1)
    counts = label_mask.sum(dim=1).tolist()  # the one host sync
    pseudo_label, pseudo_boxes, pseudo_quality = [], [], []
    for b in range(B):
        m = final_mask_sorted[b]
        if counts[b]:
            pseudo_label.append(sem_cls_label[b][m])
            pseudo_boxes.append(box_label[b][m])
            pseudo_quality.append(quality_score[b][m])
        else:
            pseudo_label.append(torch.ones((0,), device=dev))
            pseudo_boxes.append(torch.ones((0, 7), device=dev))
            pseudo_quality.append(torch.ones((0, 6), device=dev))
    return pseudo_label, pseudo_boxes, pseudo_quality
