"""ctypes binding of oracle/_ref/libnesie_ref_ops.so: the REFERENCE's own hot-path kernels
(/root/reference/mmdet3d/ops/*/src/*_cuda.cu) compiled UNMODIFIED for sm_100a by oracle/Makefile.

TEST INFRASTRUCTURE ONLY.  The launchers are C++ functions, bound by their mangled names with
raw device pointers and the current stream.  Runs only where a GPU is present (the GPU box);
used to pin the C restatement bit-for-bit, to produce tests/golden/ vectors, and as the
"reference kernels on the same B200" timing in bench.py.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libnesie_ref_ops.so")

_SYMS = {
    "fps": "_Z39furthest_point_sampling_kernel_launcheriiiPKfPfPiP11CUstream_st",
    "fps_with_dist": "_Z49furthest_point_sampling_with_dist_kernel_launcheriiiPKfPfPiP11CUstream_st",
    "ball_query": "_Z26ball_query_kernel_launcheriiiffiPKfS0_PiP11CUstream_st",
    "gather": "_Z29gather_points_kernel_launcheriiiiPKfPKiPfP11CUstream_st",
    "gather_grad": "_Z34gather_points_grad_kernel_launcheriiiiPKfPKiPfP11CUstream_st",
    "group": "_Z28group_points_kernel_launcheriiiiiPKfPKiPfP11CUstream_st",
    "group_grad": "_Z33group_points_grad_kernel_launcheriiiiiPKfPKiPfP11CUstream_st",
    "three_nn": "_Z24three_nn_kernel_launcheriiiPKfS0_PfPiP11CUstream_st",
    "three_interpolate": "_Z33three_interpolate_kernel_launcheriiiiPKfPKiS0_PfP11CUstream_st",
    "three_interpolate_grad": "_Z38three_interpolate_grad_kernel_launcheriiiiPKfPKiS0_PfP11CUstream_st",
}
_lib = None


def available():
    return os.path.exists(LIB_PATH) and torch.cuda.is_available()


def _fn(name):
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
    f = getattr(_lib, _SYMS[name])
    f.restype = None
    return f


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def furthest_point_sample(xyz, m):
    B, N, _ = xyz.shape
    idx = torch.zeros((B, m), dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    _fn("fps")(B, N, m, _p(xyz), _p(temp), _p(idx), _s())
    return idx


def furthest_point_sample_with_dist(dist, m):
    B, N, _ = dist.shape
    idx = torch.zeros((B, m), dtype=torch.int32, device=dist.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=dist.device)
    _fn("fps_with_dist")(B, N, m, _p(dist), _p(temp), _p(idx), _s())
    return idx


def ball_query(min_radius, max_radius, sample_num, xyz, center_xyz):
    B, N, _ = xyz.shape
    M = center_xyz.shape[1]
    idx = torch.zeros((B, M, sample_num), dtype=torch.int32, device=xyz.device)
    _fn("ball_query")(B, N, M, ctypes.c_float(min_radius), ctypes.c_float(max_radius), sample_num,
                      _p(center_xyz), _p(xyz), _p(idx), _s())
    return idx


def gather_points(features, indices):
    B, C, N = features.shape
    M = indices.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=features.device)
    _fn("gather")(B, C, N, M, _p(features), _p(indices), _p(out), _s())
    return out


def gather_points_grad(grad_out, indices, N):
    B, C, M = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    _fn("gather_grad")(B, C, N, M, _p(grad_out), _p(indices), _p(g), _s())
    return g


def grouping_operation(features, indices):
    B, C, N = features.shape
    _, M, K = indices.shape
    out = torch.empty((B, C, M, K), dtype=torch.float32, device=features.device)
    _fn("group")(B, C, N, M, K, _p(features), _p(indices), _p(out), _s())
    return out


def grouping_operation_grad(grad_out, indices, N):
    B, C, M, K = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    _fn("group_grad")(B, C, N, M, K, _p(grad_out), _p(indices), _p(g), _s())
    return g


def three_nn_dist2(target, source):
    B, n, _ = target.shape
    m = source.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=target.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=target.device)
    _fn("three_nn")(B, n, m, _p(target), _p(source), _p(dist2), _p(idx), _s())
    return dist2, idx


def three_nn(target, source):
    dist2, idx = three_nn_dist2(target, source)
    return torch.sqrt(dist2), idx


def three_interpolate(features, indices, weight):
    B, C, m = features.shape
    n = indices.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32, device=features.device)
    _fn("three_interpolate")(B, C, m, n, _p(features), _p(indices), _p(weight), _p(out), _s())
    return out


def three_interpolate_grad(grad_out, indices, weight, m):
    B, C, n = grad_out.shape
    g = torch.zeros((B, C, m), dtype=torch.float32, device=grad_out.device)
    _fn("three_interpolate_grad")(B, C, n, m, _p(grad_out), _p(indices), _p(weight), _p(g), _s())
    return g


# ---- points_in_boxes (roiaware_pool3d/src/points_in_boxes_cuda.cu), library built by `make ref_pib`
PIB_LIB_PATH = os.path.join(_HERE, "_ref", "libnesie_ref_pib.so")
_pib = None


def pib_available():
    return os.path.exists(PIB_LIB_PATH) and torch.cuda.is_available()


def _pib_fn(name):
    global _pib
    if _pib is None:
        _pib = ctypes.CDLL(PIB_LIB_PATH)      # resolves libtorch through its rpath
    f = getattr(_pib, name)
    f.restype = None
    return f


def points_in_boxes_gpu(points, boxes):
    """Reference launcher (default stream): (B, M, 3), (B, T, 7) -> (B, M) int32, -1 = none."""
    B, M, _ = points.shape
    out = torch.full((B, M), -1, dtype=torch.int32, device=points.device)
    torch.cuda.synchronize()
    _pib_fn("_Z24points_in_boxes_launcheriiiPKfS0_Pi")(B, boxes.shape[1], M, _p(boxes.contiguous()),
                                                      _p(points.contiguous()), _p(out))
    torch.cuda.synchronize()
    return out


def points_in_boxes_batch(points, boxes):
    B, M, _ = points.shape
    T = boxes.shape[1]
    out = torch.zeros((B, M, T), dtype=torch.int32, device=points.device)
    torch.cuda.synchronize()
    _pib_fn("_Z30points_in_boxes_batch_launcheriiiPKfS0_Pi")(B, T, M, _p(boxes.contiguous()),
                                                            _p(points.contiguous()), _p(out))
    torch.cuda.synchronize()
    return out


SORTV_PATH = os.path.join(_HERE, "_ref", "libnesie_ref_sortv.so")
_sortv = None


def sortv_available():
    return os.path.exists(SORTV_PATH) and torch.cuda.is_available()


def sort_vertices(vertices, mask, num_valid):
    """The reference's sort_vertices_wrapper (ops/rotated_iou/cuda_op/sort_vert_kernel.cu:136-139),
    which launches on the legacy default stream: synchronise around it."""
    global _sortv
    if _sortv is None:
        _sortv = ctypes.CDLL(SORTV_PATH)
    fn = getattr(_sortv, "_Z21sort_vertices_wrapperiiiPKfPKbPKiPi")
    fn.restype = None
    vertices = vertices.contiguous().float()
    mask = mask.contiguous().bool()
    num_valid = num_valid.contiguous().int()
    B, N, M, _ = vertices.shape
    idx = torch.zeros((B, N, 9), dtype=torch.int32, device=vertices.device)
    torch.cuda.synchronize()
    fn(B, N, M, _p(vertices), _p(mask), _p(num_valid), _p(idx))
    torch.cuda.synchronize()
    return idx
