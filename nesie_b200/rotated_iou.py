"""Differentiable IoU of boxes rotated about the upright axis: `cal_iou_3d` and the `sort_vertices` op.

Mirror of mmdet3d/ops/rotated_iou (oriented_iou_loss.py:86-109 cal_iou_3d, :38-58 cal_iou,
box_intersection_2d.py, cuda_op/sort_vert_kernel.cu) -- SURVEY.md 8f-2.  The IoU of NesieHead's
losses is taken between predicted boxes (whose yaw is atan2 of two regressed channels, i.e. not 0
even on ScanNet) and targets, so the polygon-clipping formulation is kept: candidate vertices =
corners of either rectangle inside the other + pairwise edge intersections (24 slots), ordered
anticlockwise by `nesie_sort_vertices` (one thread per polygon, no host sync, on the current stream),
area by the shoelace formula.  Everything except the ordering is plain differentiable tensor algebra
on (B, N, ...) tensors; gradients reach the boxes through the vertex coordinates exactly as in the
reference (the ordering itself is piecewise constant).
"""
import torch
from torch.autograd import Function

from . import _lib

EPSILON = 1e-8
_CORNER_SIGNS = {}


def _corner_signs(ref):
    """(+-0.5 corner pattern x, y) on ref's device; cached so that no host -> device copy happens
    inside a CUDA-graph capture (the first, eager call creates them)."""
    key = (ref.device, ref.dtype)
    if key not in _CORNER_SIGNS:
        _CORNER_SIGNS[key] = (torch.tensor([0.5, -0.5, -0.5, 0.5], dtype=ref.dtype, device=ref.device),
                              torch.tensor([0.5, 0.5, -0.5, -0.5], dtype=ref.dtype, device=ref.device))
    return _CORNER_SIGNS[key]


class _SortVertices(Function):

    @staticmethod
    def forward(ctx, vertices, mask, num_valid):
        _lib.need_cuda(vertices, mask, num_valid)
        vertices = vertices.contiguous().float()
        mask_u8 = mask.contiguous().to(torch.uint8)
        num_valid = num_valid.contiguous().to(torch.int32)
        B, N, M, _ = vertices.shape
        idx = torch.empty((B, N, 9), dtype=torch.int32, device=vertices.device)
        with torch.cuda.device(vertices.device):
            _lib.call("nesie_sort_vertices", B, N, M, _lib.ptr(vertices), _lib.ptr(mask_u8),
                      _lib.ptr(num_valid), _lib.ptr(idx), _lib.stream())
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad):
        return None, None, None


def sort_vertices(vertices, mask, num_valid):
    """vertices (B, N, 24, 2) fp32 centred on their mean, mask (B, N, 24) bool, num_valid (B, N)
    int32 -> (B, N, 9) int32: valid vertices in anticlockwise order, the first repeated to close the
    polygon, padded with the index of an unused intersection slot (cuda_op/cuda_ext.py:6-18)."""
    return _SortVertices.apply(vertices, mask, num_valid)


def box2corners(box):
    """(B, N, 5) x, y, w, h, alpha -> (B, N, 4, 2) corners (oriented_iou_loss.py:6-35)."""
    x, y, w, h, alpha = box.unbind(-1)
    sx, sy = _corner_signs(box)
    lx = sx * w.unsqueeze(-1)
    ly = sy * h.unsqueeze(-1)
    c, s = torch.cos(alpha).unsqueeze(-1), torch.sin(alpha).unsqueeze(-1)
    return torch.stack([lx * c - ly * s + x.unsqueeze(-1), lx * s + ly * c + y.unsqueeze(-1)], dim=-1)


def _edge_intersections(c1, c2):
    """Pairwise intersections of the 4 edges of each rectangle: ((B,N,4,4,2), mask (B,N,4,4))
    (box_intersection_2d.py:13-53; collinear edges do not intersect)."""
    a1, b1 = c1, c1.roll(-1, dims=2)                  # edge i of box 1: a1[i] -> b1[i]
    a2, b2 = c2, c2.roll(-1, dims=2)
    x1, y1 = a1[..., 0].unsqueeze(3), a1[..., 1].unsqueeze(3)
    x2, y2 = b1[..., 0].unsqueeze(3), b1[..., 1].unsqueeze(3)
    x3, y3 = a2[..., 0].unsqueeze(2), a2[..., 1].unsqueeze(2)
    x4, y4 = b2[..., 0].unsqueeze(2), b2[..., 1].unsqueeze(2)
    num = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    den_t = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)
    den_u = (x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)
    par = num == 0.0
    t = torch.where(par, torch.full_like(num, -1.0), den_t / num)
    u = torch.where(par, torch.full_like(num, -1.0), -den_u / num)
    mask = (t > 0) & (t < 1) & (u > 0) & (u < 1)
    t = den_t / (num + EPSILON)
    pts = torch.stack([x1 + t * (x2 - x1), y1 + t * (y2 - y1)], dim=-1)
    return pts * mask.float().unsqueeze(-1), mask


def _corners_inside(c1, c2):
    """Which corners of rectangle 1 lie in (or on) rectangle 2 (box_intersection_2d.py:56-83)."""
    a, b, d = c2[:, :, 0:1], c2[:, :, 1:2], c2[:, :, 3:4]
    ab, ad, am = b - a, d - a, c1 - a
    p_ab = (ab * am).sum(-1) / (ab * ab).sum(-1)
    p_ad = (ad * am).sum(-1) / (ad * ad).sum(-1)
    return (p_ab > -1e-6) & (p_ab < 1 + 1e-6) & (p_ad > -1e-6) & (p_ad < 1 + 1e-6)


def oriented_box_intersection_2d(c1, c2, sort_fn=sort_vertices):
    """Intersection area of rectangles given by corners (B, N, 4, 2) (box_intersection_2d.py:170-184)."""
    B, N = c1.shape[:2]
    inters, mask_inter = _edge_intersections(c1, c2)
    vertices = torch.cat([c1, c2, inters.reshape(B, N, 16, 2)], dim=2)
    mask = torch.cat([_corners_inside(c1, c2), _corners_inside(c2, c1), mask_inter.reshape(B, N, 16)], dim=2)
    num_valid = mask.int().sum(dim=2).int()
    mean = (vertices * mask.float().unsqueeze(-1)).sum(dim=2, keepdim=True) / num_valid[..., None, None]
    order = sort_fn(vertices - mean, mask, num_valid).long()
    sel = torch.gather(vertices, 2, order.unsqueeze(-1).expand(-1, -1, -1, 2))
    cross = sel[:, :, :-1, 0] * sel[:, :, 1:, 1] - sel[:, :, :-1, 1] * sel[:, :, 1:, 0]
    return cross.sum(dim=2).abs() / 2


class _IoU3D(Function):
    """cal_iou_3d as ONE kernel: forward value and d iou / d box1 in the same pass (nesie_iou3d)."""

    @staticmethod
    def forward(ctx, box1, box2):
        _lib.need_cuda(box1, box2)
        shape = box1.shape[:-1]
        b1 = box1.detach().reshape(-1, 7).contiguous().float()
        b2 = box2.detach().reshape(-1, 7).contiguous().float()
        n = b1.shape[0]
        iou = torch.empty((n,), dtype=torch.float32, device=b1.device)
        jac = torch.empty((n, 7), dtype=torch.float32, device=b1.device) if box1.requires_grad else None
        with torch.cuda.device(b1.device):
            _lib.call("nesie_iou3d", n, _lib.ptr(b1), _lib.ptr(b2), _lib.ptr(iou), _lib.ptr(jac),
                      _lib.stream())
        ctx.save_for_backward(jac)
        ctx.box_shape = box1.shape
        return iou.view(shape)

    @staticmethod
    def backward(ctx, grad):
        (jac,) = ctx.saved_tensors
        if jac is None:
            return None, None
        return (grad.reshape(-1, 1) * jac).view(ctx.box_shape), None


def cal_iou_3d(box3d1, box3d2, sort_fn=sort_vertices):
    """(B, N, 7) x, y, z, w, h, l, alpha boxes -> (B, N) IoU (oriented_iou_loss.py:86-109).  On CUDA
    tensors with the default vertex ordering this is the fused kernel (gradient to box3d1 only: the
    second box is a target in every loss of the path); `cal_iou_3d_tensor` is the tensor formulation."""
    if (sort_fn is sort_vertices and box3d1.is_cuda and not box3d2.requires_grad
            and box3d1.shape[-1] == 7 and box3d1.shape == box3d2.shape):
        return _IoU3D.apply(box3d1, box3d2)
    return cal_iou_3d_tensor(box3d1, box3d2, sort_fn)


def cal_iou_3d_tensor(box3d1, box3d2, sort_fn=sort_vertices):
    """The reference's tensor formulation (differentiable with respect to both boxes)."""
    bev = lambda b: torch.cat([b[..., 0:2], b[..., 3:5], b[..., 6:7]], dim=-1)   # noqa: E731  x, y, w, h, alpha
    b1, b2 = bev(box3d1), bev(box3d2)
    zmax1, zmin1 = box3d1[..., 2] + box3d1[..., 5] * 0.5, box3d1[..., 2] - box3d1[..., 5] * 0.5
    zmax2, zmin2 = box3d2[..., 2] + box3d2[..., 5] * 0.5, box3d2[..., 2] - box3d2[..., 5] * 0.5
    z_overlap = (torch.min(zmax1, zmax2) - torch.max(zmin1, zmin2)).clamp_min(0.)
    inter_area = oriented_box_intersection_2d(box2corners(b1), box2corners(b2), sort_fn)
    u = b1[..., 2] * b1[..., 3] + b2[..., 2] * b2[..., 3] - inter_area
    iou_2d = inter_area / u
    intersection_3d = iou_2d * u * z_overlap
    v1 = box3d1[..., 3] * box3d1[..., 4] * box3d1[..., 5]
    v2 = box3d2[..., 3] * box3d2[..., 4] * box3d2[..., 5]
    return intersection_3d / (v1 + v2 - intersection_3d)
