"""nesie_b200: the Nesie / VoteNet data-parallel hot path, hand-written for B200 (sm_100a).

The names exported here are the reference's `mmdet3d.ops` operator API for this path
(mmdet3d/ops/__init__.py:5-18) plus `aligned_3d_nms` (mmdet3d/core/post_processing), so the
package drops into PointSAModule / PointFPModule call sites unchanged.  Every op is a thin
ctypes call into libnesie_b200.so (C ABI: include/nesie_b200.h); nothing falls back to torch
or the CPU.
"""
from .ball_query import ball_query
from .box3d_nms import aligned_3d_nms, aligned_3d_nms_batched
from .furthest_point_sample import (Points_Sampler, calc_square_dist, furthest_point_sample,
                                    furthest_point_sample_with_dist)
from .gather_points import gather_points
from .group_points import GroupAll, QueryAndGroup, grouping_operation
from .interpolate import three_interpolate, three_nn
from .pointnet2_sa_ssg import PointNet2SASSG
from .pointnet_modules import (ConvModule, PointFPModule, PointSAModule, PointSAModuleMSG,
                               build_sa_module)
from .pseudo_label import get_pseudo_labels, lhs_3d_faster_samecls, lhs_3d_faster_samecls_batched
from .side_loss import bbox2surface, side_uncertainty_loss
from .side_pooling import MiniPointNet, SidePooling
from .points_in_boxes import points_in_boxes_batch, points_in_boxes_gpu
from .teacher_ema import TeacherEMA
from .nesie_head import NesieHead, ReliableConvBboxHead, VoteModule
from .detectors import BoxAug, VoteNet, VoteNetNesie, transformation_bbox_preds
from .ddp import FlatGradDDP
from .rotated_iou import cal_iou_3d, sort_vertices

__all__ = [
    'ball_query', 'aligned_3d_nms', 'aligned_3d_nms_batched', 'Points_Sampler',
    'calc_square_dist', 'furthest_point_sample', 'furthest_point_sample_with_dist',
    'gather_points', 'GroupAll', 'QueryAndGroup', 'grouping_operation', 'three_interpolate',
    'three_nn', 'PointNet2SASSG', 'ConvModule', 'PointFPModule', 'PointSAModule',
    'PointSAModuleMSG', 'build_sa_module', 'get_pseudo_labels', 'lhs_3d_faster_samecls',
    'lhs_3d_faster_samecls_batched', 'bbox2surface', 'side_uncertainty_loss', 'TeacherEMA',
    'SidePooling', 'MiniPointNet', 'points_in_boxes_gpu', 'points_in_boxes_batch',
    'NesieHead', 'ReliableConvBboxHead', 'VoteModule', 'VoteNet', 'VoteNetNesie', 'BoxAug',
    'transformation_bbox_preds', 'FlatGradDDP', 'cal_iou_3d', 'sort_vertices',
]
