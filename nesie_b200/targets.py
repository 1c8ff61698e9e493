"""Target assignment of the train step on the device.

Batched, padded, sync-free restatement of NesieHead.get_targets / get_targets_single
(reference: mmdet3d/models/dense_heads/nesie_head.py:511-679).  The reference runs a python loop
per scene, a python loop per GT box (`nonzero` host syncs) and chamfer_distance with (B, N, M, 3)
expands; here GT arrives padded to (B, G, 7) + a validity mask, the vote slots come from
`nesie_vote_targets`, both chamfer argmins from `nesie_chamfer_assign`, and nothing touches the host,
so the whole loss is CUDA-graph capturable.

Padding rules that reproduce the reference's per-scene lists exactly:
  * valid boxes come first in every scene, padded rows are zero (a zero box contains no point);
  * a scene without boxes is given ONE fake zero box with label 0 and valid mask 0
    (nesie_head.py:531-541): slot 0 of its padded row plays that part, so its proposals are
    assigned to the origin, exactly as in the reference;
  * centre targets are zero-padded to the batch's largest box count (:568-574) and the centre
    loss's source->target minimum runs over those padded (0, 0, 0) targets too: slots beyond that
    count are excluded on the device.
"""
import torch

from . import _lib


def vote_targets(points, boxes, n_valid, seed_indices=None):
    """points (B, N, >=3) fp32, boxes (B, G, 7) depth frame bottom-centred, n_valid (B,) int32.
    -> (vote_targets (B, R, 9) fp32, vote_target_masks (B, R) int64) with R = N, or the seeds'
    rows only when seed_indices (B, S) int64 is given."""
    _lib.need_cuda(points, boxes)
    points = points.contiguous().float()
    boxes = boxes.contiguous().float()
    B, N, C = points.shape
    G = boxes.shape[1]
    idx = None
    R = N
    if seed_indices is not None:
        idx = seed_indices.contiguous().long()
        R = idx.shape[1]
    vt = torch.empty((B, R, 9), dtype=torch.float32, device=points.device)
    vm = torch.empty((B, R), dtype=torch.int64, device=points.device)
    n_valid = n_valid.contiguous().to(torch.int32)
    with torch.cuda.device(points.device):
        _lib.call("nesie_vote_targets", B, N, G, R, _lib.ptr(points), C, _lib.ptr(boxes),
                  _lib.ptr(n_valid), _lib.ptr(idx), _lib.ptr(vt), _lib.ptr(vm), _lib.stream())
    return vt, vm


def chamfer_assign(src, dst, n_valid=None, want=(True, True)):
    """src (B, N, 3), dst (B, M, 3) -> (idx1 (B, N) int64 nearest dst slot among the first n_valid[b],
    idx2 (B, M) int64 nearest src point), squared-L2 criterion, first minimum on ties."""
    _lib.need_cuda(src, dst)
    src = src.detach().contiguous().float()
    dst = dst.detach().contiguous().float()
    B, N, _ = src.shape
    M = dst.shape[1]
    idx1 = torch.empty((B, N), dtype=torch.int64, device=src.device) if want[0] else None
    idx2 = torch.empty((B, M), dtype=torch.int64, device=src.device) if want[1] else None
    if n_valid is not None:
        n_valid = n_valid.contiguous().to(torch.int32)
    with torch.cuda.device(src.device):
        _lib.call("nesie_chamfer_assign", B, N, M, _lib.ptr(src), _lib.ptr(dst), _lib.ptr(n_valid),
                  _lib.ptr(idx1), _lib.ptr(idx2), _lib.stream())
    return idx1, idx2


def pad_gt(gt_boxes, gt_labels, device, pad_to=None, extra=None):
    """Per-scene lists ((n_i, 7) tensors or objects with `.tensor`, (n_i,) labels) -> padded
    (B, G, 7) boxes, (B, G) int64 labels, (B, G) bool validity [, (B, G, E) padded `extra` rows]."""
    tensors = [b.tensor if hasattr(b, "tensor") else b for b in gt_boxes]
    B = len(tensors)
    G = max([t.shape[0] for t in tensors] + [1])
    if pad_to is not None:
        assert pad_to >= G, "pad_to smaller than the largest box count"
        G = pad_to
    boxes = torch.zeros((B, G, 7), dtype=torch.float32)
    labels = torch.zeros((B, G), dtype=torch.int64)
    valid = torch.zeros((B, G), dtype=torch.bool)
    ex = None
    if extra is not None:
        E = max([e.shape[-1] for e in extra if e.dim() == 2] + [1])
        ex = torch.zeros((B, G, E), dtype=torch.float32)
    for i, t in enumerate(tensors):
        n = t.shape[0]
        if n:
            boxes[i, :n] = t.detach().float().cpu()[:, :7]
            labels[i, :n] = gt_labels[i].detach().cpu().long()
            valid[i, :n] = True
            if extra is not None:
                ex[i, :n] = extra[i].detach().float().cpu()
    out = (boxes.to(device), labels.to(device), valid.to(device))
    return out + (ex.to(device),) if extra is not None else out


def gravity_center(boxes):
    """(.., 7) bottom-centred depth boxes -> (.., 3) gravity centres (depth_box3d.py:42-48)."""
    gc = boxes[..., :3].clone()
    gc[..., 2] = boxes[..., 2] + boxes[..., 5] * 0.5
    return gc
