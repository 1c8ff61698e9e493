"""ctypes binding of libnesie_b200.so (the C ABI declared in include/nesie_b200.h).

The shared library is the product: there is no Python/torch fallback for any op.  If the
library is missing, or a tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnesie_b200.so")

_i, _f, _d, _p, _ll = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_void_p, ctypes.c_longlong

# name -> argtypes, exactly the prototypes of include/nesie_b200.h
SIGNATURES = {
    "nesie_abi_version": [],
    "nesie_fps": [_i, _i, _i, _p, _p, _p, _p],
    "nesie_fps_needs_temp": [_i, _i, _i],
    "nesie_fps_with_dist": [_i, _i, _i, _p, _p, _p, _p],
    "nesie_ball_query": [_i, _i, _i, _f, _f, _i, _p, _p, _p, _p],
    "nesie_ball_query_grid_workspace": [_i, _i, _i],
    "nesie_ball_query_grid": [_i, _i, _i, _f, _f, _i, _p, _p, _p, _p, _ll, _p],
    "nesie_gather_points": [_i, _i, _i, _i, _p, _p, _p, _p],
    "nesie_gather_points_grad": [_i, _i, _i, _i, _p, _p, _p, _p],
    "nesie_group_points": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "nesie_group_points_grad": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "nesie_query_group_concat": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p],
    "nesie_group_rows": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _i, _p],
    "nesie_group_rows_grad": [_i, _i, _i, _i, _i, _p, _p, _f, _p, _p, _p, _i, _p],
    "nesie_three_nn": [_i, _i, _i, _p, _p, _p, _p, _p],
    "nesie_three_interpolate": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "nesie_three_interpolate_grad": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "nesie_points_in_boxes": [_i, _i, _i, _p, _p, _p, _p],
    "nesie_points_in_boxes_batch": [_i, _i, _i, _p, _p, _p, _p],
    "nesie_group_max_rows_forward": [_ll, _i, _i, _p, _p, _p, _p, _i, _p],
    "nesie_group_max_bias_parts": [_ll, _i],
    "nesie_group_max_rows_backward": [_ll, _i, _i, _p, _p, _p, _i, _p, _p],
    "nesie_interp_rows": [_i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p],
    "nesie_aligned_3d_nms_batched": [_i, _i, _p, _p, _p, _p, _f, _p, _p, _p],
    "nesie_lhs_nms_batched": [_i, _i, _p, _p, _d, _i, _p, _p, _p],
    "nesie_side_uncertainty_loss": [_i, _i, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p],
    "nesie_side_uncertainty_loss_grad": [_i, _i, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p],
    "nesie_ema_update": [_ll, _p, _p, _f, _f, _p],
    "nesie_gemm_b_image_bytes": [_i, _i],
    "nesie_gemm_pack_b": [_i, _i, _ll, _ll, _p, _p, _p],
    "nesie_gemm_nt_3xtf32": [_ll, _i, _i, _p, _ll, _p, _p, _ll, _p],
    "nesie_gemm_debug_profile": [_p],
    "nesie_gemm_fused_supported": [_ll, _i, _i, _p, _ll, _ll],
    "nesie_gemm_stats_parts": [_ll],
    "nesie_gemm_nt_3xtf32_fused": [_ll, _i, _i, _p, _ll, _p, _p, _ll, _p, _p, _p, _p],
    "nesie_gemm_nt_3xtf32_bnbwd": [_ll, _i, _i, _p, _ll, _p, _p, _ll, _p, _ll, _p, _p, _p],
    "nesie_gemm_nt_3xtf32_pool": [_ll, _i, _i, _p, _ll, _p, _p, _ll, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _p],
    "nesie_three_nn_grid": [_i, _i, _i, _p, _p, _p, _p, _p, _ll, _i, _p],
    "nesie_gather_linear_parts": [_i, _i, _i],
    "nesie_gather_linear_forward": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _f, _p, _p, _p],
    "nesie_gather_linear_backward": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _f, _p, _p, _p],
    "nesie_pool_finalize": [_ll, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "nesie_bn_pool_finalize": [_ll, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "nesie_pool_wgrad_parts": [_ll],
    "nesie_pool_wgrad": [_ll, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "nesie_pool_dgrad": [_ll, _i, _i, _i, _p, _p, _p, _p, _p],
    "nesie_group_sum_rows": [_ll, _i, _i, _p, _p, _p],
    "nesie_colsum_rows_workspace": [_i],
    "nesie_colsum_rows": [_ll, _i, _p, _p, _p, _p],
    "nesie_scatter_rows_add": [_ll, _i, _i, _p, _p, _p, _p],
    "nesie_bn_relu_rows_backward_fused": [_ll, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p],
    "nesie_gemm_wgrad_3xtf32_fused": [_ll, _i, _i, _p, _ll, _p, _ll, _p, _p, _p, _i, _p],
    "nesie_bn_rows_forward_fused": [_ll, _i, _i, _p, _p, _p, _f, _f, _p, _p, _p, _i, _p, _p, _p, _p, _p],
    "nesie_gemm_sum_partials": [_i, _ll, _p, _p, _p],
    "nesie_gemm_wgrad_splits": [_ll, _i, _i],
    "nesie_gemm_wgrad_3xtf32": [_ll, _i, _i, _p, _ll, _p, _ll, _p, _i, _p],
    "nesie_bn_rows_workspace_bytes": [_i],
    "nesie_bn_relu_rows_forward": [_ll, _i, _i, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    "nesie_bn_relu_rows_backward": [_ll, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "nesie_sa_fused_supported": [_i, _i, _i, _i, _i],
    "nesie_pack_features_bf16": [_i, _i, _i, _p, _p, _p],
    "nesie_vote_targets": [_i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p],
    "nesie_chamfer_assign": [_i, _i, _i, _p, _p, _p, _p, _p, _p],
    "nesie_sort_vertices": [_i, _i, _i, _p, _p, _p, _p, _p],
    "nesie_iou3d": [_ll, _p, _p, _p, _p, _p],
    "nesie_sa_fused_forward": [_i] * 8 + [_p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p],
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m nesie_b200.build` "
                "(there is no CPU or torch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _ll if name.endswith(('_workspace', '_bytes')) else _i
        handle.nesie_last_error.restype = ctypes.c_char_p
        handle.nesie_last_error.argtypes = []
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().nesie_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nesie_b200 ops run on CUDA tensors only (no CPU fallback)")


LAUNCHES = 0  # kernels launched through the C ABI by this process (every entry point = 1 launch)


TRACE = None  # when a list: every call appends (name, args) -- bench.py's per-shape kernel census


def call(name, *args):
    global LAUNCHES
    LAUNCHES += 1
    if TRACE is not None:
        TRACE.append((name, args))
    check(getattr(lib(), name)(*args), name)
