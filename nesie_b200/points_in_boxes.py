"""points_in_boxes_gpu / points_in_boxes_batch with the reference's python interface
(mmdet3d/ops/roiaware_pool3d/points_in_boxes.py:6-48,84-123): same arguments, asserts, output
dtype / initial values; used by the box structures' `points_in_boxes` for target assignment
(core/bbox/structures/depth_box3d.py:251-277, nesie_head.py:628,749)."""
import torch

from . import _lib


def _check(points, boxes):
    assert boxes.shape[0] == points.shape[0], \
        f'Points and boxes should have the same batch size, got {boxes.shape[0]} and {points.shape[0]}'
    assert boxes.shape[2] == 7, f'boxes dimension should be 7, got unexpected shape {boxes.shape[2]}'
    assert points.shape[2] == 3, f'points dimension should be 3, got unexpected shape {points.shape[2]}'
    _lib.need_cuda(points, boxes)
    assert points.device == boxes.device, 'Points and boxes should be put on the same device'


def points_in_boxes_gpu(points, boxes):
    """points (B, M, 3), boxes (B, T, 7) [x, y, z (bottom centre), w, l, h, ry] in LiDAR coordinates
    -> (B, M) int32 index of the first box containing each point, -1 for background."""
    _check(points, boxes)
    B, M, _ = points.shape
    out = points.new_zeros((B, M), dtype=torch.int).fill_(-1)
    with torch.cuda.device(points.device):
        boxes_f, points_f = boxes.float().contiguous(), points.float().contiguous()   # alive across the launch
        _lib.call("nesie_points_in_boxes", B, boxes.shape[1], M, _lib.ptr(boxes_f),
                  _lib.ptr(points_f), _lib.ptr(out), _lib.stream())
    return out


def points_in_boxes_batch(points, boxes):
    """-> (B, M, T) int32, 1 where point m lies inside box t (boxes may overlap), 0 elsewhere."""
    _check(points, boxes)
    B, M, _ = points.shape
    T = boxes.shape[1]
    # every (point, box) flag is written by the kernel: no zero fill of the 4*B*M*T byte output
    out = torch.empty((B, M, T), dtype=torch.int, device=points.device)
    with torch.cuda.device(points.device):
        boxes_f, points_f = boxes.float().contiguous(), points.float().contiguous()   # alive across the launch
        _lib.call("nesie_points_in_boxes_batch", B, T, M, _lib.ptr(boxes_f),
                  _lib.ptr(points_f), _lib.ptr(out), _lib.stream())
    return out
