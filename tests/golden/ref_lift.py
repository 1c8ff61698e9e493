"""Runs the REFERENCE's own python source on the CPU of the build container (golden generation only).

`import mmdet3d` is impossible here (mmcv / mmdet / mmseg are absent and cannot be installed), so the
classes and functions of the hot path are lifted out of their files under /root/reference with `ast`
and exec'd UNMODIFIED into one namespace that supplies:
  * the third-party pieces they import, restated from the pinned versions of env_setup.sh:3-6
    (mmdet==2.19.0 `weighted_loss` / `weight_reduce_loss` / MSELoss / L1Loss / SmoothL1Loss /
    CrossEntropyLoss of mmdet/models/losses/{utils,mse_loss,smooth_l1_loss,cross_entropy_loss}.py,
    `multi_apply` of mmdet/core/utils/misc.py; mmcv-full==1.3.17 ConvModule / build_conv_layer /
    BaseModule / force_fp32 / Hook / registries),
  * the CUDA extensions, supplied by the C oracle (oracle/nesie_oracle.c: bit-identical to the
    reference kernels, tests/test_oracle_cpu.py) -- furthest_point_sample, points_in_boxes_batch,
    the sort_vertices op of ops/rotated_iou, and PointSAModule through oracle/modules.py,
  * `Tensor.cuda()` / `.to('cuda')` as no-ops.
Nothing here is imported by the product or by the GPU-box tests; only the make_golden_*.py scripts
next to it use it, and they store inputs + outputs as small .npz fixtures.
"""
import ast
import functools
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF = "/root/reference/mmdet3d"


# ---------------------------------------------------------------------------------------------
# third-party restatements (mmdet 2.19.0 / mmcv-full 1.3.17)
# ---------------------------------------------------------------------------------------------
def reduce_loss(loss, reduction):
    """mmdet/models/losses/utils.py::reduce_loss"""
    e = F._Reduction.get_enum(reduction)
    if e == 0:
        return loss
    if e == 1:
        return loss.mean()
    return loss.sum()


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """mmdet/models/losses/utils.py::weight_reduce_loss"""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        loss = reduce_loss(loss, reduction)
    elif reduction == 'mean':
        loss = loss.sum() / avg_factor
    elif reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def weighted_loss(loss_func):
    """mmdet/models/losses/utils.py::weighted_loss"""
    @functools.wraps(loss_func)
    def wrapper(pred, target, weight=None, reduction='mean', avg_factor=None, **kwargs):
        loss = loss_func(pred, target, **kwargs)
        return weight_reduce_loss(loss, weight, reduction, avg_factor)
    return wrapper


@weighted_loss
def _mse_loss(pred, target):
    return F.mse_loss(pred, target, reduction='none')


@weighted_loss
def _l1_loss(pred, target):
    if target.numel() == 0:
        return pred.sum() * 0
    assert pred.size() == target.size()
    return torch.abs(pred - target)


@weighted_loss
def _smooth_l1_loss(pred, target, beta=1.0):
    assert beta > 0
    if target.numel() == 0:
        return pred.sum() * 0
    diff = torch.abs(pred - target)
    return torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta)


class MSELoss(nn.Module):
    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * _mse_loss(pred, target, weight, reduction=reduction,
                                            avg_factor=avg_factor)


class L1Loss(nn.Module):
    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * _l1_loss(pred, target, weight, reduction=reduction,
                                           avg_factor=avg_factor)


class SmoothL1Loss(nn.Module):
    def __init__(self, beta=1.0, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.beta, self.reduction, self.loss_weight = beta, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kw):
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * _smooth_l1_loss(pred, target, weight, beta=self.beta,
                                                  reduction=reduction, avg_factor=avg_factor, **kw)


class CrossEntropyLoss(nn.Module):
    """mmdet/models/losses/cross_entropy_loss.py (softmax branch: use_sigmoid=False, use_mask=False)"""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None,
                 ignore_index=None, loss_weight=1.0):
        super().__init__()
        assert not use_sigmoid and not use_mask
        self.reduction, self.loss_weight = reduction, loss_weight
        self.class_weight, self.ignore_index = class_weight, ignore_index

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None,
                ignore_index=None, **kwargs):
        reduction = reduction_override if reduction_override else self.reduction
        if ignore_index is None:
            ignore_index = self.ignore_index
        ignore_index = -100 if ignore_index is None else ignore_index
        cw = None
        if self.class_weight is not None:
            cw = cls_score.new_tensor(self.class_weight, device=cls_score.device)
        loss = F.cross_entropy(cls_score, label, weight=cw, reduction='none',
                               ignore_index=ignore_index)
        if weight is not None:
            weight = weight.float()
        return self.loss_weight * weight_reduce_loss(loss, weight=weight, reduction=reduction,
                                                     avg_factor=avg_factor)


def multi_apply(func, *args, **kwargs):
    """mmdet/core/utils/misc.py::multi_apply"""
    pfunc = functools.partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def force_fp32(*a, **k):
    return lambda f: f


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


class Hook:
    pass


def is_tuple_of(seq, expected_type):
    return isinstance(seq, tuple) and all(isinstance(s, expected_type) for s in seq)


def build_conv_layer(cfg, *args, **kwargs):
    """mmcv/cnn/bricks/conv.py::build_conv_layer for the two types the path uses."""
    layer = {None: nn.Conv2d, 'Conv1d': nn.Conv1d, 'Conv2d': nn.Conv2d}[cfg['type'] if cfg else None]
    return layer(*args, **kwargs)


# ---------------------------------------------------------------------------------------------
# lifting
# ---------------------------------------------------------------------------------------------
def lift(relpath, ns, names=None, consts=True):
    """exec the top-level class / function definitions (and simple NAME = constant assignments) of a
    reference file into `ns`, unmodified.  `names`: restrict to these definitions."""
    path = relpath if os.path.isabs(relpath) else os.path.join(REF, relpath)
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)):
            if names is None or node.name in names:
                exec(compile(ast.Module([node], []), path, "exec"), ns)
        elif consts and isinstance(node, ast.Assign) and len(node.targets) == 1 and \
                isinstance(node.targets[0], ast.Name) and isinstance(node.value, (ast.Constant, ast.UnaryOp)):
            if names is None or node.targets[0].id in names:
                exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def patch_cuda_noop():
    torch.Tensor.cuda = lambda self, *a, **k: self
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(torch.device('cpu') if (isinstance(x, str) and x.startswith('cuda')) else x for x in a)
        return _to(self, *a, **k)
    torch.Tensor.to = to


def sort_v(vertices, mask, num_valid):
    """ops/rotated_iou/cuda_op/sort_vert_kernel.cu through the C oracle restatement."""
    from oracle import cpu
    return cpu.sort_vertices(vertices, mask, num_valid)


def base_namespace():
    """Namespace with the stubs; reference definitions are lifted into it in dependency order."""
    from oracle import cpu
    import nesie_b200 as nb
    from oracle import modules as om

    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "numpy": np, "functools": functools,
          "weighted_loss": weighted_loss, "MSELoss": MSELoss, "L1Loss": L1Loss,
          "SmoothL1Loss": SmoothL1Loss, "multi_apply": multi_apply, "force_fp32": force_fp32,
          "BaseModule": BaseModule, "Hook": Hook, "is_tuple_of": is_tuple_of,
          "build_conv_layer": build_conv_layer, "ConvModule": nb.ConvModule,
          "HEADS": _Registry(), "LOSSES": _Registry(), "DETECTORS": _Registry(), "HOOKS": _Registry(),
          "is_module_wrapper": lambda m: False, "sort_v": sort_v,
          "l1_loss": F.l1_loss, "mse_loss": F.mse_loss, "smooth_l1_loss": F.smooth_l1_loss,
          "furthest_point_sample": cpu.furthest_point_sample,
          "three_nn": lambda target, source: cpu.three_nn(target, source),
          "points_in_boxes_batch": lambda pts, boxes: cpu.points_in_boxes_batch(pts, boxes),
          "aligned_3d_nms": None, "cal_giou_3d": None, "iou3d_cuda": None, "BasePoints": (),
          "Counter": __import__("collections").Counter, "abstractmethod": lambda f: f}
    ns["mmcv"] = types.SimpleNamespace()
    # `from .box_3d_mode import Box3DMode` inside DepthInstance3DBoxes.points_in_boxes: a stub package
    ns["__name__"] = "mmdet3d.core.bbox.structures.depth_box3d"
    ns["__package__"] = "mmdet3d.core.bbox.structures"
    pkg = ""
    for part in "mmdet3d.core.bbox.structures".split("."):
        pkg = part if not pkg else pkg + "." + part
        m = sys.modules.setdefault(pkg, types.ModuleType(pkg))
        m.__path__ = []
    mode = types.ModuleType("mmdet3d.core.bbox.structures.box_3d_mode")
    mode.Box3DMode = types.SimpleNamespace(LIDAR=0, CAM=1, DEPTH=2)
    sys.modules[mode.__name__] = mode

    # box structures (core/bbox/structures): utils -> base -> mode -> depth
    lift("core/bbox/structures/utils.py", ns, names={"limit_period", "rotation_3d_in_axis", "xywhr2xyxyr"})
    lift("core/bbox/structures/base_box3d.py", ns)
    lift("core/bbox/structures/depth_box3d.py", ns)
    ns["_DepthBoxes"] = ns["DepthInstance3DBoxes"]

    class DepthInstance3DBoxes(ns["_DepthBoxes"]):
        """convert_to(LIDAR) of the reference goes through Box3DMode.convert
        (core/bbox/structures/box_3d_mode.py:124-160): xyz @ [[0,1,0],[-1,0,0],[0,0,1]]^T, sizes
        (y, x, z), remaining columns unchanged, returned as a LiDAR box (only `.tensor` is read)."""

        def convert_to(self, dst, rt_mat=None):
            arr = self.tensor.clone()
            x_size, y_size, z_size = arr[..., 3:4], arr[..., 4:5], arr[..., 5:6]
            rt = arr.new_tensor([[0, 1, 0], [-1, 0, 0], [0, 0, 1]])
            xyz = arr[:, :3] @ rt.t()
            out = torch.cat([xyz[:, :3], torch.cat([y_size, x_size, z_size], dim=-1), arr[:, 6:]], dim=-1)
            return types.SimpleNamespace(tensor=out)
    ns["DepthInstance3DBoxes"] = DepthInstance3DBoxes

    # losses
    lift("models/losses/chamfer_distance.py", ns)
    ns["CrossEntropyLoss"] = CrossEntropyLoss
    lift("models/losses/surface_loss.py", ns)
    lift("models/losses/side_pred_loss.py", ns)
    lift("models/losses/gfocal_loss.py", ns)
    lift("ops/rotated_iou/box_intersection_2d.py", ns)
    lift("ops/rotated_iou/oriented_iou_loss.py", ns, names={"box2corners_th", "cal_iou", "cal_iou_3d"})
    lift("models/losses/iou3d_loss.py", ns, names={"iou_3d_loss", "IoU3DMixin", "IoU3DLoss"})

    def build_loss(cfg):
        cfg = dict(cfg)
        return ns[cfg.pop("type")](**cfg)
    ns["build_loss"] = build_loss

    # modules
    lift("models/model_utils/vote_module.py", ns)
    lift("models/dense_heads/reliable_conv_bbox_module.py", ns)
    lift("models/dense_heads/side_pooling_module.py", ns)

    def build_sa_module(cfg, *a, **k):
        """PointSAModule parameters in the reference layout; forward on the CPU oracle ops."""
        cfg = dict(cfg)
        assert cfg.pop("type") == "PointSAModule"
        mod = nb.PointSAModule(**cfg)
        mod.forward = lambda points_xyz, features=None, indices=None, target_xyz=None: \
            om.sa_forward(mod, points_xyz, features, indices, target_xyz)
        return mod
    ns["build_sa_module"] = build_sa_module
    lift("models/dense_heads/nesie_head.py", ns)
    return ns


NESIE_HEAD_CFG = dict(   # configs/Nesie/nesie-votenet-scannet-train-010.py:17-93
    num_classes=18, reg_max=32, alpha=1.0,
    vote_module_cfg=dict(in_channels=256, vote_per_seed=1, gt_per_seed=3, conv_channels=(256, 256),
                         conv_cfg=dict(type="Conv1d"), norm_cfg=dict(type="BN1d"), norm_feats=True,
                         vote_loss=dict(type="ChamferDistance", mode="l1", reduction="none",
                                        loss_dst_weight=10.0)),
    vote_aggregation_cfg=dict(type="PointSAModule", num_point=256, radius=0.3, num_sample=16,
                              mlp_channels=[256, 128, 128, 128], use_xyz=True, normalize_xyz=True),
    pred_layer_cfg=dict(in_channels=128, shared_conv_channels=(128, 128), bias=True),
    objectness_loss=dict(type="CrossEntropyLoss", class_weight=[0.2, 0.8], reduction="sum", loss_weight=5.0),
    center_loss=dict(type="ChamferDistance", mode="l2", reduction="sum", loss_src_weight=10.0,
                     loss_dst_weight=10.0),
    iou_loss=dict(type="IoU3DLoss", reduction="sum", loss_weight=3.0),
    semantic_loss=dict(type="CrossEntropyLoss", reduction="sum", loss_weight=1.0),
    iou_pred_loss=dict(type="GeneralQualityFocalLoss", reduction="sum", use_sigmoid=False, beta=2.0,
                       loss_weight=1.0),
    surface_loss=dict(type="SurfaceLoss", func_type="MSELoss", beta=5.0, reduction="sum", loss_weight=10.0),
    side_loss=dict(type="SidePredLoss", label_func_type="SmoothL1Loss", loss_func_type="MSELoss",
                   beta=5.0, reduction="sum", loss_weight=1.0),
    train_cfg=dict(pos_distance_thr=0.3, neg_distance_thr=0.6, sample_mod="vote",
                   dataset_name="ScanNet", thresh_warmup=True, use_cbl=True),
)
