"""ctypes binding of libnesie_oracle.so (oracle/nesie_oracle.c) over CPU torch tensors.

TEST INFRASTRUCTURE ONLY.  Function names and argument meaning follow the reference's python
ops (mmdet3d/ops/*), so parity tests read like the reference's own call sites.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnesie_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "nesie_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libnesie_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.nesie_oracle_opt_n_threads.restype = ctypes.c_int
    return _lib


def set_threads(n):
    lib().nesie_oracle_set_threads(int(n))


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    return t.detach().cpu().contiguous().float()


def _i32(t):
    return t.detach().cpu().contiguous().to(torch.int32)


def opt_n_threads(n):
    return lib().nesie_oracle_opt_n_threads(int(n))


def furthest_point_sample(xyz, m):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, dtype=torch.float32)
    idx = torch.zeros((B, m), dtype=torch.int32)
    lib().nesie_oracle_fps(B, N, m, _p(xyz), _p(temp), _p(idx))
    return idx


def furthest_point_sample_with_dist(dist, m):
    dist = _f32(dist)
    B, N, _ = dist.shape
    temp = torch.full((B, N), 1e10, dtype=torch.float32)
    idx = torch.zeros((B, m), dtype=torch.int32)
    lib().nesie_oracle_fps_with_dist(B, N, m, _p(dist), _p(temp), _p(idx))
    return idx


def ball_query(min_radius, max_radius, sample_num, xyz, center_xyz):
    xyz, center_xyz = _f32(xyz), _f32(center_xyz)
    B, N, _ = xyz.shape
    M = center_xyz.shape[1]
    idx = torch.zeros((B, M, sample_num), dtype=torch.int32)
    lib().nesie_oracle_ball_query(B, N, M, ctypes.c_float(min_radius), ctypes.c_float(max_radius),
                                  sample_num, _p(center_xyz), _p(xyz), _p(idx))
    return idx


def gather_points(features, indices):
    features, indices = _f32(features), _i32(indices)
    B, C, N = features.shape
    M = indices.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32)
    lib().nesie_oracle_gather_points(B, C, N, M, _p(features), _p(indices), _p(out))
    return out


def gather_points_grad(grad_out, indices, N):
    grad_out, indices = _f32(grad_out), _i32(indices)
    B, C, M = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32)
    lib().nesie_oracle_gather_points_grad(B, C, N, M, _p(grad_out), _p(indices), _p(g))
    return g


def grouping_operation(features, indices):
    features, indices = _f32(features), _i32(indices)
    B, C, N = features.shape
    _, M, K = indices.shape
    out = torch.empty((B, C, M, K), dtype=torch.float32)
    lib().nesie_oracle_group_points(B, C, N, M, K, _p(features), _p(indices), _p(out))
    return out


def grouping_operation_grad(grad_out, indices, N):
    grad_out, indices = _f32(grad_out), _i32(indices)
    B, C, M, K = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32)
    lib().nesie_oracle_group_points_grad(B, C, N, M, K, _p(grad_out), _p(indices), _p(g))
    return g


def three_nn_dist2(target, source):
    """Raw kernel outputs: (squared distances, idx)."""
    target, source = _f32(target), _f32(source)
    B, n, _ = target.shape
    m = source.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32)
    idx = torch.empty((B, n, 3), dtype=torch.int32)
    lib().nesie_oracle_three_nn(B, n, m, _p(target), _p(source), _p(dist2), _p(idx))
    return dist2, idx


def three_nn(target, source):
    """Returns (sqrt(dist2), idx) like the reference's python wrapper (three_nn.py:38)."""
    dist2, idx = three_nn_dist2(target, source)
    return torch.sqrt(dist2), idx


def three_interpolate(features, indices, weight):
    features, indices, weight = _f32(features), _i32(indices), _f32(weight)
    B, C, m = features.shape
    n = indices.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32)
    lib().nesie_oracle_three_interpolate(B, C, m, n, _p(features), _p(indices), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, indices, weight, m):
    grad_out, indices, weight = _f32(grad_out), _i32(indices), _f32(weight)
    B, C, n = grad_out.shape
    g = torch.zeros((B, C, m), dtype=torch.float32)
    lib().nesie_oracle_three_interpolate_grad(B, C, n, m, _p(grad_out), _p(indices), _p(weight),
                                              _p(g))
    return g


def points_in_boxes_gpu(points, boxes):
    """(B, M, 3), (B, T, 7) -> (B, M) int32 index of the first containing box, -1 for none
    (the arithmetic of the reference's CUDA kernel, run on the CPU)."""
    points, boxes = _f32(points), _f32(boxes)
    B, M, _ = points.shape
    T = boxes.shape[1]
    out = torch.full((B, M), -1, dtype=torch.int32)
    lib().nesie_oracle_points_in_boxes(B, T, M, _p(boxes), _p(points), _p(out), 0)
    return out


def points_in_boxes_batch(points, boxes):
    """(B, M, 3), (B, T, 7) -> (B, M, T) int32 membership mask."""
    points, boxes = _f32(points), _f32(boxes)
    B, M, _ = points.shape
    T = boxes.shape[1]
    out = torch.zeros((B, M, T), dtype=torch.int32)
    lib().nesie_oracle_points_in_boxes(B, T, M, _p(boxes), _p(points), _p(out), 1)
    return out


def sort_vertices(vertices, mask, num_valid):
    """ops/rotated_iou/cuda_op (sort_vertices_forward): vertices (B, N, 24, 2) normalised around
    their mean, mask (B, N, 24) bool, num_valid (B, N) int32 -> (B, N, 9) int32 anticlockwise
    vertex order, first index repeated, padded with an invalid intersection slot."""
    v = _f32(vertices)
    mk = mask.detach().cpu().contiguous().to(torch.uint8)
    nv = _i32(num_valid)
    B, N, M, _ = v.shape
    idx = torch.zeros((B, N, 9), dtype=torch.int32)
    lib().nesie_oracle_sort_vertices(B, N, M, _p(v), _p(mk), _p(nv), _p(idx))
    return idx
