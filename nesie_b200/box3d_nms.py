"""aligned_3d_nms: drop-in for mmdet3d/core/post_processing/box3d_nms.py:129-176, plus the batched
form the B200 path uses (all scenes of a step in ONE launch, no host sync per picked box)."""
import torch

from . import _lib


def aligned_3d_nms_batched(boxes, scores, classes, thresh, counts=None):
    """boxes (S, n, 6) [x1,y1,z1,x2,y2,z2], scores (S, n), classes (S, n), counts (S,) or None.

    Returns (keep (S, n) int64, keep_cnt (S,) int32): keep[s, :keep_cnt[s]] are the picked box
    indices of scene s in pick order (descending score); the remaining slots are -1.
    Nothing is synchronised with the host."""
    _lib.need_cuda(boxes, scores, classes)
    S, n = scores.shape
    boxes = boxes.contiguous().float()
    scores = scores.contiguous().float()
    classes = classes.contiguous().to(torch.int32)
    if counts is not None:
        counts = counts.contiguous().to(torch.int32)
    keep = torch.full((S, n), -1, dtype=torch.int64, device=boxes.device)
    keep_cnt = torch.zeros((S,), dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        _lib.call("nesie_aligned_3d_nms_batched", S, n, _lib.ptr(boxes), _lib.ptr(scores),
                  _lib.ptr(classes), _lib.ptr(counts), float(thresh), _lib.ptr(keep),
                  _lib.ptr(keep_cnt), _lib.stream())
    return keep, keep_cnt


def aligned_3d_nms(boxes, scores, classes, thresh):
    """Single-scene API of the reference: boxes (n, 6), scores (n,), classes (n,) ->
    (k,) int64 indices of the selected boxes, highest score first."""
    if boxes.shape[0] == 0:
        return boxes.new_zeros((0,), dtype=torch.long)
    keep, cnt = aligned_3d_nms_batched(boxes[None], scores[None], classes[None], thresh)
    return keep[0, :int(cnt.item())]
