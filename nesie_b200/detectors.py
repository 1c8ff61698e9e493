"""VoteNet (supervised pretrain step) and VoteNetNesie (mean-teacher step) on this repo's kernels.

Mirror of mmdet3d/models/detectors/votenet.py:27-60 (VoteNet.forward_train) and
votenet_nesie.py:36-127,301-324,596-634 (combine/choose_{sup,unsup}_item, forward_train, ulb_update,
transformation_bbox_preds).  Same order of operations as the reference step:

    student forward on every scene (labeled + unlabeled, strong view)
    no_grad:  switch_to_teacher (EMA weights swapped in, model stays in train(): batch-stat BN)
              teacher forward on every scene (weak view) -> get_pseudo_labels
              pseudo boxes: teacher frame -> original frame -> student frame
              switch_to_student
    supervised losses on the labeled rows, ulb_update, unsupervised losses on the unlabeled rows

What changes is where it runs.  The reference pulls the teacher's predictions to the host (>= 7
.cpu() copies), builds pseudo-label lists in python, moves the boxes to the CPU to transform them and
loops over scenes / GT boxes for the targets.  Here the whole step stays on the device with static
shapes: pseudo labels stay packed as (B, 64, .) + a mask (pseudo_label.get_pseudo_labels(as_lists=
False)) and are compacted by a stable sort, the augmentation flow is applied to all boxes of all scenes
at once (`BoxAug`), the class-count tables (`ulb_list`, `ulb_flag`: SimiRunnerHook state) are device
buffers updated by scatter, and the losses take padded targets (nesie_head.py) -- so the step, the
optimizer and the EMA update can be captured in ONE CUDA graph.
"""
import math

import torch
from torch import nn as nn

from . import bn_rows
from .nesie_head import NesieHead
from .pointnet2_sa_ssg import PointNet2SASSG
from .pseudo_label import MAX_NUM_OBJ, get_pseudo_labels
from .teacher_ema import TeacherEMA

NUM_CLASSES = 18


def nesie_head_cfg(num_classes=NUM_CLASSES, mean_size_arr_path=None, num_proposal=256, **over):
    """NesieHead arguments of configs/Nesie/nesie-votenet-scannet-train-010.py:17-93."""
    cfg = dict(
        num_classes=num_classes, reg_max=32, alpha=1.0,
        vote_module_cfg=dict(in_channels=256, vote_per_seed=1, gt_per_seed=3, conv_channels=(256, 256),
                             conv_cfg=dict(type='Conv1d'), norm_cfg=dict(type='BN1d'), norm_feats=True,
                             vote_loss=dict(type='ChamferDistance', mode='l1', reduction='none',
                                            loss_dst_weight=10.0)),
        vote_aggregation_cfg=dict(type='PointSAModule', num_point=num_proposal, radius=0.3, num_sample=16,
                                  mlp_channels=[256, 128, 128, 128], use_xyz=True, normalize_xyz=True),
        pred_layer_cfg=dict(in_channels=128, shared_conv_channels=(128, 128), bias=True),
        objectness_loss=dict(type='CrossEntropyLoss', class_weight=[0.2, 0.8], reduction='sum',
                             loss_weight=5.0),
        center_loss=dict(type='ChamferDistance', mode='l2', reduction='sum', loss_src_weight=10.0,
                         loss_dst_weight=10.0),
        iou_loss=dict(type='IoU3DLoss', reduction='sum', loss_weight=3.0),
        semantic_loss=dict(type='CrossEntropyLoss', reduction='sum', loss_weight=1.0),
        iou_pred_loss=dict(type='GeneralQualityFocalLoss', reduction='sum', use_sigmoid=False, beta=2.0,
                           loss_weight=1.0),
        surface_loss=dict(type='SurfaceLoss', func_type='MSELoss', beta=5.0, reduction='sum',
                          loss_weight=10.0),
        side_loss=dict(type='SidePredLoss', label_func_type='SmoothL1Loss', loss_func_type='MSELoss',
                       beta=5.0, reduction='sum', loss_weight=1.0),
        grid_conv_cfg=dict(num_class=num_classes, num_heading_bin=1, num_size_cluster=num_classes,
                           mean_size_arr_path=mean_size_arr_path, num_proposal=num_proposal,
                           sampling='seed_fps', query_feats='seed'),
        train_cfg=dict(pos_distance_thr=0.3, neg_distance_thr=0.6, sample_mod='vote',
                       dataset_name='ScanNet', thresh_warmup=True, use_cbl=True))
    cfg.update(over)
    return cfg


class VoteNet(nn.Module):
    """backbone -> bbox_head -> bbox_head.loss (detectors/votenet.py:27-60)."""

    head_cls = NesieHead
    backbone_cls = PointNet2SASSG

    def __init__(self, backbone=None, bbox_head=None, train_cfg=None):
        super().__init__()
        self.backbone = self.backbone_cls(**(backbone or dict(in_channels=4)))
        bbox_head = dict(bbox_head or nesie_head_cfg())
        self.train_cfg = dict(train_cfg or bbox_head.get('train_cfg') or {})
        bbox_head['train_cfg'] = self.train_cfg
        self.bbox_head = self.head_cls(**bbox_head)

    def extract_feat(self, points, fps_indices=None, after_level=None):
        if fps_indices is not None or after_level is not None:
            return self.backbone(points, fps_indices=fps_indices, after_level=after_level)
        return self.backbone(points)

    def predict(self, points, fps_indices=None, after_level=None, jitter_noise=None):
        with bn_rows.defer_batch_counters():
            x = self.extract_feat(points, fps_indices, after_level)
            return self.bbox_head(x, self.train_cfg.get('sample_mod', 'vote'),
                                  self.train_cfg.get('dataset_name', 'ScanNet'), jitter_noise=jitter_noise)

    def forward(self, points, **kw):
        return self.predict(points, **kw)

    def forward_train(self, points, gt_bboxes_3d, gt_labels_3d, **kw):
        """points (B, N, 4) or list; per-scene GT lists -> loss dict (list interface)."""
        pts = torch.stack(list(points)) if isinstance(points, (list, tuple)) else points
        return self.bbox_head.loss(self.predict(pts, **kw), pts, gt_bboxes_3d, gt_labels_3d)

    def forward_train_padded(self, points, boxes, labels, valid, **kw):
        """Static-shape form: GT padded to (B, G, 7) / (B, G) / (B, G) bool (CUDA-graph capturable)."""
        return self.bbox_head.loss_padded(self.predict(points, **kw), points, boxes, labels, valid)


class BoxAug:
    """The augmentation record of a batch of scenes (img_metas of the reference: RandomFlip3D and
    GlobalRotScaleTrans write `transformation_3d_flow` = [HF] [VF] R S T, `pcd_rotation`,
    `pcd_scale_factor`, `pcd_trans`) as batched device tensors."""

    def __init__(self, hf, vf, rot, scale, trans):
        self.hf, self.vf, self.rot, self.scale, self.trans = hf, vf, rot, scale, trans

    @classmethod
    def identity(cls, B, device):
        return cls(torch.zeros(B, dtype=torch.bool, device=device), torch.zeros(B, dtype=torch.bool, device=device),
                   torch.eye(3, device=device).repeat(B, 1, 1), torch.ones(B, device=device),
                   torch.zeros(B, 3, device=device))

    @classmethod
    def random(cls, B, device, generator=None, rot_range=0.087266, scale_range=(1.0, 1.0),
               trans_std=0.0):
        """One draw of the reference's train pipeline (flip each BEV axis with p = 0.5, rotation in
        +-5 degrees; configs/Nesie/*-train-010.py data pipeline)."""
        r = lambda *s: torch.rand(*s, generator=generator)   # noqa: E731
        ang = (r(B) * 2 - 1) * rot_range
        c, s = torch.cos(ang), torch.sin(ang)
        rot = torch.zeros(B, 3, 3)
        rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1], rot[:, 2, 2] = c, s, -s, c, 1.0
        scale = scale_range[0] + r(B) * (scale_range[1] - scale_range[0])
        trans = torch.randn(B, 3, generator=generator) * trans_std
        return cls((r(B) < 0.5).to(device), (r(B) < 0.5).to(device), rot.to(device), scale.to(device),
                   trans.to(device))

    def index(self, idx):
        return BoxAug(self.hf[idx], self.vf[idx], self.rot[idx], self.scale[idx], self.trans[idx])

    def apply_points(self, points):
        """Augment (B, N, >=3) points the way the pipeline does (flip, rotate, scale, translate)."""
        p = points.clone()
        p[..., 0] = torch.where(self.hf[:, None], -p[..., 0], p[..., 0])
        p[..., 1] = torch.where(self.vf[:, None], -p[..., 1], p[..., 1])
        p[..., :3] = p[..., :3] @ self.rot
        p[..., :3] = p[..., :3] * self.scale[:, None, None] + self.trans[:, None, :]
        return p


def _flip(boxes, flag, axis):
    """DepthInstance3DBoxes.flip (core/bbox/structures/depth_box3d.py:176-195) where flag is set."""
    f = flag[:, None]
    out = boxes.clone()
    out[..., axis] = torch.where(f, -boxes[..., axis], boxes[..., axis])
    yaw = -boxes[..., 6] + math.pi if axis == 0 else -boxes[..., 6]
    out[..., 6] = torch.where(f, yaw, boxes[..., 6])
    return out


def _rotate(boxes, rot_mat_T):
    """DepthInstance3DBoxes.rotate with a matrix (depth_box3d.py:133-153): xyz @ rot_mat_T,
    yaw -= atan2(rot_mat_T[0, 1], rot_mat_T[0, 0])."""
    out = boxes.clone()
    out[..., :3] = boxes[..., :3] @ rot_mat_T
    out[..., 6] = boxes[..., 6] - torch.atan2(rot_mat_T[:, 0, 1], rot_mat_T[:, 0, 0])[:, None]
    return out


def untransform_boxes(boxes, aug):
    """votenet_nesie.py:596-614: undo T, S, R, VF, HF (the flow reversed); boxes (B, G, 7)."""
    b = boxes.clone()
    b[..., :3] = b[..., :3] + (-1.0 * aug.trans)[:, None, :]
    b[..., :6] = b[..., :6] * (1.0 / aug.scale)[:, None, None]
    b = _rotate(b, aug.rot.transpose(1, 2))       # rotate(pcd_rotation): rot_mat_T = pcd_rotation.T
    b = _flip(b, aug.vf, 1)
    return _flip(b, aug.hf, 0)


def transform_boxes(boxes, aug):
    """votenet_nesie.py:616-634: apply HF, VF, R, S, T."""
    b = _flip(boxes, aug.hf, 0)
    b = _flip(b, aug.vf, 1)
    b = _rotate(b, aug.rot)                       # rotate(pcd_rotation.T): rot_mat_T = pcd_rotation
    b = b.clone()
    b[..., :6] = b[..., :6] * aug.scale[:, None, None]
    b[..., :3] = b[..., :3] + aug.trans[:, None, :]
    return b


def transformation_bbox_preds(boxes, aug_t=None, aug_s=None):
    """votenet_nesie.py:310-324 on padded device boxes (B, G, 7)."""
    if aug_t is not None:
        boxes = untransform_boxes(boxes, aug_t)
    if aug_s is not None:
        boxes = transform_boxes(boxes, aug_s)
    return boxes


def ulb_update(ulb_list, ulb_flag, positions, labels, valid):
    """votenet_nesie.py:301-308: for every unlabeled scene of the batch (table row positions[i]) clear
    its flag and overwrite its row with the class histogram of its pseudo labels.
    labels (Bu, G) int64, valid (Bu, G) bool.  In place, no host sync."""
    C = ulb_list.shape[1]
    hist = torch.zeros((labels.shape[0], C), dtype=ulb_list.dtype, device=ulb_list.device)
    hist.scatter_add_(1, labels.clamp(0, C - 1), valid.to(ulb_list.dtype))
    ulb_flag.index_fill_(0, positions, 0.0)
    ulb_list.index_copy_(0, positions, hist)


def choose_items(bbox_preds, index):
    """choose_sup_item / choose_unsup_item (votenet_nesie.py:46-67) with a precomputed row index."""
    return {k: v.index_select(0, index) for k, v in bbox_preds.items() if torch.is_tensor(v)}


def compact_pseudo_labels(packed):
    """Packed get_pseudo_labels output -> (boxes (B, 64, 7) bottom-centred, labels, valid, quality)
    with the selected entries first, in their original order (the order of the reference's lists)."""
    mask = packed['label_mask'].bool()
    order = torch.argsort((~mask).to(torch.int8), dim=1, stable=True)
    box = torch.cat([packed['center_label'], packed['size_label'], packed['heading_label']], dim=-1)
    g = lambda t: torch.gather(t, 1, order.unsqueeze(-1).expand(-1, -1, t.shape[-1]))   # noqa: E731
    valid = torch.gather(mask, 1, order)
    boxes = g(box) * valid.unsqueeze(-1).to(box.dtype)
    return boxes, torch.gather(packed['sem_cls_label'], 1, order), valid, g(packed['quality_score'])


class VoteNetNesie(VoteNet):
    """Mean-teacher detector (detectors/votenet_nesie.py).  `n_lb` / `n_ulb`: sizes of the labeled /
    unlabeled scene tables the reference's runner attaches (lb_map / ulb_map)."""

    def __init__(self, backbone=None, bbox_head=None, train_cfg=None, n_lb=120, n_ulb=1081,
                 ema=dict(momentum=0.001, interval=1, warm_up=10), quality_poly=(5 / 3, 8 / 3),
                 teacher_mode='overlap'):
        """teacher_mode: 'swap' -- the reference's schedule (student forward, swap EMA weights in,
        teacher forward, swap back); 'overlap' -- the teacher pass reads the EMA copies directly and
        runs on its own stream BESIDE the student forward (same results: the BatchNorm running
        statistics, the only state both passes write, are merged afterwards in the reference's
        student-then-teacher order)."""
        super().__init__(backbone, bbox_head, train_cfg)
        assert teacher_mode in ('swap', 'overlap')
        self.teacher_mode = teacher_mode
        self._overlap = None
        self.n_lb, self.n_ulb = n_lb, n_ulb
        self.quality_poly = quality_poly
        self.ema_cfg = dict(ema)
        C = self.bbox_head.num_classes
        # SimiRunnerHook.before_run state (core/utils/simi_runner_hook.py:52-71)
        self.register_buffer('ulb_list', torch.zeros(n_ulb, C))
        self.register_buffer('ulb_flag', torch.ones(n_ulb))
        self.teacher = None

    def init_teacher(self, flat=None):
        """SimiTeacherHook.hooks_before_run: EMA copies of every parameter (call after .to(device)).
        flat: a FlatGradDDP(flatten_parameters=True) whose flat parameter buffer is shared."""
        self.teacher = TeacherEMA(self, flat=flat, **self.ema_cfg)
        return self.teacher

    def _filter_teacher(self, preds_t, aug_t, aug_s):
        packed = get_pseudo_labels(
            preds_t, self.ulb_list, self.ulb_flag, self.n_lb, self.n_ulb,
            num_classes=self.bbox_head.num_classes,
            thresh_warmup=self.train_cfg.get('thresh_warmup', True),
            use_cbl=self.train_cfg.get('use_cbl', True), quality_poly=self.quality_poly,
            as_lists=False)
        boxes, labels, valid, quality = compact_pseudo_labels(packed)
        return transformation_bbox_preds(boxes, aug_t, aug_s), labels, valid, quality

    # ---- teacher pass beside the student pass -------------------------------------------------------
    def _overlap_state(self, dev):
        if self._overlap is None:
            bns = [(n, m) for n, m in self.named_modules()
                   if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.track_running_stats]
            real, names, mom = [], [], []
            for n, m in bns:
                for b in ('running_mean', 'running_var'):
                    real.append(getattr(m, b))
                    names.append(f'{n}.{b}')
                    mom.append(m.momentum)
            nbt = [m.num_batches_tracked for _, m in bns]
            self._overlap = dict(
                stream=torch.cuda.Stream(device=dev), real=real, names=names, keep=[1.0 - x for x in mom],
                neg_keep=[-(1.0 - x) for x in mom], snap=[t.clone() for t in real],
                clone=[t.clone() for t in real], nbt=nbt, nbt_names=[f'{n}.num_batches_tracked' for n, _ in bns],
                nbt_clone=[t.clone() for t in nbt])
        return self._overlap

    def _teacher_begin(self, points_t, aug_t, aug_s, **kw):
        """Fork: the teacher forward + pseudo-label filter on the teacher stream, reading the EMA copies
        through torch.func.functional_call and writing BatchNorm statistics into private clones."""
        from torch.func import functional_call
        dev = points_t.device
        st = self._overlap_state(dev)
        main = torch.cuda.current_stream(dev)
        torch._foreach_copy_(st['snap'], st['real'])
        torch._foreach_copy_(st['clone'], st['real'])
        torch._foreach_copy_(st['nbt_clone'], st['nbt'])
        tensors = dict(self.teacher.ema_named_views())
        tensors.update(zip(st['names'], st['clone']))
        tensors.update(zip(st['nbt_names'], st['nbt_clone']))
        st['stream'].wait_stream(main)
        cross = [points_t] + [t for a in (aug_t, aug_s) if a is not None
                              for t in (a.hf, a.vf, a.rot, a.scale, a.trans)]
        cross += [t for v in kw.values() if isinstance(v, (list, tuple)) for t in v if torch.is_tensor(t)]
        for t in cross:
            t.record_stream(st['stream'])
        with torch.cuda.stream(st['stream']), torch.no_grad():
            preds_t = functional_call(self, tensors, (points_t,), kw)
            out = self._filter_teacher(preds_t, aug_t, aug_s)
        return out

    def _teacher_end(self, dev):
        """Join, then merge the running statistics as if the teacher pass had run after the student
        pass: r <- (1 - m) r_student + (r_teacher_clone - (1 - m) r_before)."""
        st = self._overlap
        torch.cuda.current_stream(dev).wait_stream(st['stream'])
        torch._foreach_mul_(st['real'], st['keep'])
        torch._foreach_add_(st['real'], st['clone'])
        torch._foreach_mul_(st['snap'], st['neg_keep'])
        torch._foreach_add_(st['real'], st['snap'])
        torch._foreach_add_(st['nbt'], 1)

    def teacher_pseudo_labels(self, points_t, aug_t=None, aug_s=None, **kw):
        """no_grad teacher pass -> padded pseudo boxes in the student frame."""
        with torch.no_grad():
            self.teacher.swap()
            preds_t = self.predict(points_t, **kw)
            packed = get_pseudo_labels(
                preds_t, self.ulb_list, self.ulb_flag, self.n_lb, self.n_ulb,
                num_classes=self.bbox_head.num_classes,
                thresh_warmup=self.train_cfg.get('thresh_warmup', True),
                use_cbl=self.train_cfg.get('use_cbl', True), quality_poly=self.quality_poly,
                as_lists=False)
            boxes, labels, valid, quality = compact_pseudo_labels(packed)
            boxes = transformation_bbox_preds(boxes, aug_t, aug_s)
            self.teacher.swap()
        return boxes, labels, valid, quality

    def forward_train_padded(self, points_s, points_t, gt_boxes, gt_labels, gt_valid, sup_index,
                             unsup_index, ulb_positions, aug_t=None, aug_s=None,
                             student_kw=None, teacher_kw=None):
        """points_s / points_t (B, N, 4): student / teacher views of every scene; gt_* padded GT of the
        labeled rows (len(sup_index), G, .); sup_index / unsup_index: int64 row indices of the labeled /
        unlabeled scenes; ulb_positions (len(unsup_index),) int64 rows of the unlabeled scenes in the
        class-count table."""
        if self.teacher_mode == 'overlap' and points_t.is_cuda:
            pl_boxes, pl_labels, pl_valid, pl_quality = self._teacher_begin(
                points_t, aug_t, aug_s, **(teacher_kw or {}))
            preds_s = self.predict(points_s, **(student_kw or {}))
            self._teacher_end(points_t.device)
        else:
            preds_s = self.predict(points_s, **(student_kw or {}))
            pl_boxes, pl_labels, pl_valid, pl_quality = self.teacher_pseudo_labels(
                points_t, aug_t, aug_s, **(teacher_kw or {}))
        head = self.bbox_head
        sup = head.loss_padded(choose_items(preds_s, sup_index), points_s.index_select(0, sup_index),
                               gt_boxes, gt_labels, gt_valid)
        u = lambda t: t.index_select(0, unsup_index)   # noqa: E731
        ulb_update(self.ulb_list, self.ulb_flag, ulb_positions, u(pl_labels), u(pl_valid))
        unsup = head.unsup_loss_padded(choose_items(preds_s, unsup_index), u(points_s), u(pl_boxes),
                                       u(pl_labels), u(pl_valid), u(pl_quality))
        return {**sup, **unsup}

    def after_train_iter(self, curr_step):
        """SimiRunnerHook.after_train_iter -> SimiTeacherHook.hooks_after_train_iter."""
        self.teacher.after_train_iter(curr_step)


class VoteNetSAQE(VoteNetNesie):
    """SAQE variant (detectors/votenet_saqe.py differs from votenet_nesie.py at :121,170,201): quality
    polynomial 0.8 s^2 - 1.8 s + 1 in the pseudo-label filter and the SAQE head's uncertainty weighting
    of the surface / IoU terms (dense_heads/saqe_head.py:590-607,631-641)."""

    def __init__(self, backbone=None, bbox_head=None, **kw):
        bbox_head = dict(bbox_head or nesie_head_cfg())
        bbox_head.setdefault('uncertainty', 'saqe')
        kw.setdefault('quality_poly', (0.8, 1.8))
        super().__init__(backbone, bbox_head, **kw)
