"""NesieHead: vote -> aggregate -> predict -> side2box -> SidePooling quality head, with the
reference's losses and target assignment, on this repo's kernels.

Mirror of mmdet3d/models/dense_heads/nesie_head.py (forward :211-275, loss :277-412, unsup_loss
:414-509, get_targets :511-591, get_targets_single :593-679), of ReliableConvBboxHead
(dense_heads/reliable_conv_bbox_module.py) and VoteModule (model_utils/vote_module.py:85-185): same
constructor arguments, sub-module / state_dict names (`vote_module.*`, `vote_aggregation.*`,
`conv_pred.shared_convs.layer{i}.*`, `conv_pred.conv_{cls,bbox,heading}.*`, `integral.project`,
`grid_conv.*`), prediction dict keys and loss dict keys.

What changes is the evaluation:
  * every 1x1 Conv1d stack runs as row GEMMs on the tcgen05 3xTF32 kernels (conv_rows.py), the three
    prediction convs as one GEMM over their concatenated weights;
  * the per-row python gathers (`torch.stack([side_pred[i, :, indx[i]] ...])`, nesie_head.py:346,386)
    are tensor gathers, the surface loss with per-side uncertainty is one fused kernel each way
    (side_loss.py);
  * target assignment is batched and sync-free (targets.py: `nesie_vote_targets`,
    `nesie_chamfer_assign`); GT arrives as per-scene lists like in the reference (`loss`,
    `unsup_loss`) or pre-padded (`loss_padded`, `unsup_loss_padded`: static shapes, CUDA-graph
    capturable);
  * cal_iou_3d's vertex ordering is `nesie_sort_vertices` on the current stream (rotated_iou.py).
One deliberate difference: IoU3DLoss returns early through a host-side `torch.any(weight > 0)`
(models/losses/iou3d_loss.py:57-58); here the weighted sum is always formed (same value, no sync).

`uncertainty='saqe'` selects the SAQE head's weighting of the surface / IoU terms
(dense_heads/saqe_head.py:590-607,631-641: exp(-sigma.detach()), no alpha * sigma term).
Hooks prefixed `_k_` are the kernels; oracle/nesie_head_ref.py overrides exactly those with CPU
restatements (test infrastructure).
"""
import os

import torch
from torch import nn as nn
from torch.nn import functional as F

from . import targets as T
from .branches import run_branches
from .conv_rows import conv1d_rows
from .furthest_point_sample import furthest_point_sample
from .pointnet_modules import ConvModule, build_sa_module, _rows_linear
from .rotated_iou import cal_iou_3d, sort_vertices
from .side_loss import bbox2surface, side_uncertainty_loss


class Integral(nn.Module):
    """sum_i P(y_i) * y_i over the reg_max + 1 bins of a side distance (nesie_head.py:19-52)."""

    def __init__(self, reg_max=16):
        super().__init__()
        self.reg_max = reg_max
        self.register_buffer('project', torch.linspace(0, reg_max, reg_max + 1) / reg_max)

    def forward(self, x):
        x = F.softmax(x.reshape(-1, self.reg_max + 1), dim=1)
        return F.linear(x, self.project.type_as(x)).reshape(-1, 6)


class ReliableConvBboxHead(nn.Module):
    """shared convs -> {cls, bbox, heading} 1x1 convs (reliable_conv_bbox_module.py:10-178).  Only
    the configuration the Nesie configs use is built: no per-branch conv stacks."""

    def __init__(self, in_channels=0, shared_conv_channels=(), cls_conv_channels=(),
                 num_cls_out_channels=0, bbox_conv_channels=(), num_bbox_out_channels=0,
                 heading_conv_channels=(), num_heading_out_channels=0, reg_max=16,
                 conv_cfg=dict(type='Conv1d'), norm_cfg=dict(type='BN1d'),
                 act_cfg=dict(type='ReLU'), bias='auto', init_cfg=None):
        super().__init__()
        assert in_channels > 0 and num_cls_out_channels > 0 and num_bbox_out_channels > 0
        assert num_heading_out_channels > 0
        if cls_conv_channels or bbox_conv_channels or heading_conv_channels:
            raise NotImplementedError("per-branch conv stacks are not used by any Nesie / SAQE config")
        prev = in_channels
        self.shared_conv_channels = tuple(shared_conv_channels)
        if self.shared_conv_channels:
            self.shared_convs = nn.Sequential()
            for i, c in enumerate(self.shared_conv_channels):
                self.shared_convs.add_module(f'layer{i}', ConvModule(
                    prev, c, 1, conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg, bias=bias))
                prev = c
        self.conv_cls = nn.Conv1d(prev, num_cls_out_channels, 1)
        self.conv_bbox = nn.Conv1d(prev, num_bbox_out_channels, 1)
        self.conv_heading = nn.Conv1d(prev, num_heading_out_channels, 1)

    def forward(self, feats):
        """(B, C, P) -> cls_score (B, n_cls, P), bbox_pred (B, n_bbox + n_heading, P)."""
        rows, B, P = self.forward_rows(feats)
        n_cls = self.conv_cls.out_channels
        out = rows.view(B, P, -1).transpose(1, 2)
        return out[:, :n_cls], out[:, n_cls:]

    def forward_rows(self, feats):
        """Point-major result: rows (B*P, n_cls + n_bbox + n_heading), one GEMM for the three convs."""
        B, _, P = feats.shape
        x = conv1d_rows(self.shared_convs, feats) if self.shared_conv_channels else feats
        convs = (self.conv_cls, self.conv_bbox, self.conv_heading)
        if not x.is_cuda:
            out = torch.cat([c(x) for c in convs], dim=1)
            return out.transpose(1, 2).reshape(B * P, -1), B, P
        w = torch.cat([c.weight.flatten(1) for c in convs], dim=0)
        bias = torch.cat([c.bias for c in convs], dim=0)
        r = x.transpose(1, 2).reshape(B * P, -1)
        return _rows_linear(r, w) + bias, B, P


class VoteModule(nn.Module):
    """seed (xyz, feats) -> votes (xyz + offset, feats + residual, L2-normalised) and the vote
    loss (model_utils/vote_module.py:85-185; vote_per_seed = 1, with_res_feat, no xyz range)."""

    def __init__(self, in_channels, vote_per_seed=1, gt_per_seed=3, num_points=-1,
                 conv_channels=(16, 16), conv_cfg=dict(type='Conv1d'), norm_cfg=dict(type='BN1d'),
                 act_cfg=dict(type='ReLU'), norm_feats=True, with_res_feat=True,
                 vote_xyz_range=None, vote_loss=None):
        super().__init__()
        if vote_per_seed != 1 or not with_res_feat or vote_xyz_range is not None or num_points != -1:
            raise NotImplementedError("only the VoteModule configuration of the Nesie configs is built")
        assert gt_per_seed == 3, "nesie_vote_targets fills three slots per seed"
        self.in_channels, self.vote_per_seed, self.gt_per_seed = in_channels, vote_per_seed, gt_per_seed
        self.norm_feats = norm_feats
        self.vote_loss_cfg = dict(vote_loss or dict(type='ChamferDistance', mode='l1', reduction='none',
                                                    loss_dst_weight=10.0))
        assert self.vote_loss_cfg.get('mode', 'l2') == 'l1' and self.vote_loss_cfg.get('reduction') == 'none'
        layers, prev = [], in_channels
        for c in conv_channels:
            layers.append(ConvModule(prev, c, 1, conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg,
                                     bias=True))
            prev = c
        self.vote_conv = nn.Sequential(*layers)
        self.conv_out = nn.Conv1d(prev, (3 + in_channels) * vote_per_seed, 1)

    def forward(self, seed_points, seed_feats):
        votes = conv1d_rows(self.conv_out, conv1d_rows(self.vote_conv, seed_feats)).transpose(2, 1)
        offset = votes[..., 0:3]
        vote_points = (seed_points + offset).contiguous()
        vote_feats = (seed_feats.transpose(2, 1) + votes[..., 3:]).transpose(2, 1).contiguous()
        if self.norm_feats:
            vote_feats = vote_feats.div(torch.norm(vote_feats, p=2, dim=1).unsqueeze(1))
        return vote_points, vote_feats, offset.transpose(2, 1)

    def get_loss(self, seed_points, vote_points, seed_indices, vote_targets_mask, vote_targets,
                 at_seeds=False):
        """vote_module.py:149-185.  at_seeds: the targets are already the seeds' rows."""
        B, S = seed_points.shape[:2]
        if at_seeds:
            mask, gt = vote_targets_mask.float(), vote_targets
        else:
            mask = torch.gather(vote_targets_mask, 1, seed_indices).float()
            gt = torch.gather(vote_targets, 1, seed_indices.unsqueeze(-1).expand(-1, -1, 9))
        gt = gt + seed_points.repeat(1, 1, self.gt_per_seed)
        weight = mask / (torch.sum(mask) + 1e-6)
        # ChamferDistance(l1, 'none') target->source term with ONE source point per seed
        dist = (vote_points.view(B * S, 1, 3) - gt.view(B * S, 3, 3)).abs().sum(-1)
        dist = dist * weight.view(B * S, 1) * self.vote_loss_cfg.get('loss_dst_weight', 1.0)
        return torch.sum(torch.min(dist, dim=1)[0])


def quality_focal_loss(pred, label, score, weight, beta=2.0, loss_weight=1.0):
    """GeneralQualityFocalLoss(use_sigmoid=False, reduction='sum') on probabilities
    (models/losses/gfocal_loss.py:8-51): every class is pushed to 0 with |p|^beta, the labelled
    class to the IoU `score` with |score - p|^beta.  Every row carries a foreground label here."""
    neg = F.binary_cross_entropy(pred, torch.zeros_like(pred), reduction='none') * pred.pow(beta)
    p_pos = torch.gather(pred, 1, label.unsqueeze(1)).squeeze(1)
    pos = F.binary_cross_entropy(p_pos, score, reduction='none') * (score - p_pos).abs().pow(beta)
    onehot = label.unsqueeze(1) == torch.arange(pred.shape[1], device=pred.device).unsqueeze(0)
    loss = torch.where(onehot, pos.unsqueeze(1), neg).sum(dim=1)
    return loss_weight * (loss * weight).sum()


class NesieHead(nn.Module):

    def __init__(self, num_classes, reg_max=16, reg_channels=128, train_cfg=None, test_cfg=None,
                 vote_module_cfg=None, vote_aggregation_cfg=None, pred_layer_cfg=None, alpha=0.5,
                 objectness_loss=None, center_loss=None, semantic_loss=None, iou_loss=None,
                 iou_pred_loss=None, surface_loss=None, side_loss=None, init_cfg=None,
                 grid_conv_cfg=None, sizes=(3.0, 3.0, 2.5), uncertainty='nesie'):
        super().__init__()
        assert uncertainty in ('nesie', 'saqe')
        self.num_classes, self.reg_max, self.reg_channels = num_classes, reg_max, reg_channels
        self.train_cfg, self.test_cfg = train_cfg or {}, test_cfg or {}
        self.gt_per_seed = vote_module_cfg['gt_per_seed']
        self.num_proposal = vote_aggregation_cfg['num_point']
        self.alpha, self.sizes, self.uncertainty = alpha, tuple(sizes), uncertainty
        # loss hyper-parameters (the reference builds mmdet loss modules from these dicts)
        self.loss_cfg = dict(
            objectness=dict(objectness_loss or dict(class_weight=[0.2, 0.8], loss_weight=5.0)),
            center=dict(center_loss or dict(loss_src_weight=10.0, loss_dst_weight=10.0)),
            semantic=dict(semantic_loss or dict(loss_weight=1.0)),
            iou=dict(iou_loss or dict(loss_weight=3.0)),
            iou_pred=dict(iou_pred_loss or dict(beta=2.0, loss_weight=1.0)),
            surface=dict(surface_loss or dict(loss_weight=10.0)),
            side=dict(side_loss or dict(loss_weight=1.0)))
        assert self.loss_cfg['surface'].get('func_type', 'MSELoss') == 'MSELoss'
        assert self.loss_cfg['center'].get('mode', 'l2') == 'l2'
        self.vote_module = VoteModule(**vote_module_cfg)
        self.vote_aggregation = build_sa_module(vote_aggregation_cfg)
        self.n_reg_outs = 6 * (reg_max + 1)
        self.conv_pred = ReliableConvBboxHead(**pred_layer_cfg, num_cls_out_channels=num_classes + 2,
                                              num_bbox_out_channels=self.n_reg_outs,
                                              num_heading_out_channels=2, reg_max=reg_max)
        self.integral = Integral(reg_max)
        # constants as non-persistent buffers (the state_dict stays the reference's; no host -> device
        # copies at run time, so the step can be captured in a CUDA graph)
        sx, sy, sz = self.sizes
        self.register_buffer('_scale_vec', torch.tensor([sx, sy, sz, sx, sy, sz]), persistent=False)
        cw = self.loss_cfg['objectness'].get('class_weight')
        self.register_buffer('_obj_class_weight', torch.tensor(cw) if cw is not None else None,
                             persistent=False)
        self.grid_conv = self._k_side_pooling_cls()(**grid_conv_cfg)

    # ---- kernels (the CPU oracle twin overrides these) ------------------------------------------
    @staticmethod
    def _k_side_pooling_cls():
        from .side_pooling import SidePooling
        return SidePooling

    def _k_aggregate(self, **kw):
        return self.vote_aggregation(**kw)

    def _k_fps(self, xyz, n):
        return furthest_point_sample(xyz, n)

    def _k_vote_targets(self, points, boxes, n_valid, seed_indices):
        return T.vote_targets(points, boxes, n_valid, seed_indices)

    def _k_chamfer_assign(self, src, dst, n_valid, want=(True, True)):
        return T.chamfer_assign(src, dst, n_valid, want)

    _k_sort_vertices = staticmethod(sort_vertices)

    def _k_side_loss(self, surface_pred, box_targets, side_scores, sem_scores, weight):
        lw = self.loss_cfg['surface'].get('loss_weight', 1.0)
        if self.uncertainty == 'saqe':
            loss, sigma = side_uncertainty_loss(surface_pred, box_targets, side_scores.detach(),
                                                sem_scores, weight, lw, 0.0)
            return loss, sigma
        return side_uncertainty_loss(surface_pred, box_targets, side_scores, sem_scores, weight, lw,
                                     self.alpha)

    # ---- forward --------------------------------------------------------------------------------
    def _extract_input(self, feat_dict):
        return feat_dict['fp_xyz'][-1], feat_dict['fp_features'][-1], feat_dict['fp_indices'][-1]

    def side2box(self, aggregated_points, reg_rows, results):
        """reg_rows (B, P, 6 * (reg_max + 1) + 2) -> surface_pred, surface_scale, bbox_preds, bbox_probs
        (nesie_head.py:150-176 + :255-258)."""
        B, P = reg_rows.shape[:2]
        prob = F.softmax(reg_rows[..., :self.n_reg_outs].reshape(B, P, 6, self.reg_max + 1), dim=3)
        res = F.linear(prob, self.integral.project.type_as(prob))                   # (B, P, 6)
        scale = self._scale_vec.to(res.dtype).expand(B, P, 6)
        d = res * scale
        lo, hi = aggregated_points - d[..., :3], aggregated_points + d[..., 3:]
        results['surface_pred'] = torch.cat([lo, hi], dim=-1)
        results['surface_scale'] = scale
        hs, hc = reg_rows[..., self.n_reg_outs], reg_rows[..., self.n_reg_outs + 1]
        norm = torch.pow(torch.pow(hs, 2) + torch.pow(hc, 2), 0.5)
        heading = torch.atan2(hs / norm, hc / norm)
        results['bbox_preds'] = torch.cat([(lo + hi) / 2.0, hi - lo, heading.unsqueeze(-1)], dim=-1)
        results['bbox_probs'] = prob.permute(0, 2, 3, 1)                            # (B, 6, bins, P)
        return results

    def jitter_bbox_preds(self, results, dataset_name, noise=None):
        """Doubles the proposals with a randomly jittered copy (nesie_head.py:178-209).  noise: the
        two N(0, 1) draws of shape (B, P, 3) (centre, size); drawn from the device generator when
        None (capturable in a CUDA graph)."""
        bp = results['bbox_preds']
        center, size, heading = bp[..., :3], bp[..., 3:6], bp[..., 6]
        if noise is None:
            noise = (torch.randn(size.shape, device=size.device), torch.randn(size.shape, device=size.device))
        center_j = center + size * noise[0] * 0.3
        size_j = torch.clamp(size + size * noise[1] * 0.3, min=1e-8)
        results['jitter_bbox_preds'] = torch.cat([center_j, size_j, heading.unsqueeze(-1)], dim=-1)
        heading2 = torch.cat([heading, heading], dim=1)
        if dataset_name == 'ScanNet':
            heading2 = torch.zeros_like(heading2)
        return torch.cat([center, center_j], dim=1), torch.cat([size, size_j], dim=1), heading2, results

    def forward(self, feat_dict, sample_mod, dataset_name='ScanNet', jitter_noise=None):
        assert sample_mod in ['vote', 'seed', 'random', 'spec']
        seed_points, seed_features, seed_indices = self._extract_input(feat_dict)
        vote_points, vote_features, vote_offset = self.vote_module(seed_points, seed_features)
        results = dict(seed_points=seed_points, seed_features=seed_features, seed_indices=seed_indices,
                       vote_points=vote_points, vote_features=vote_features, vote_offset=vote_offset)
        if sample_mod == 'vote':
            agg_in = dict(points_xyz=vote_points, features=vote_features)
        elif sample_mod == 'seed':
            agg_in = dict(points_xyz=vote_points, features=vote_features,
                          indices=self._k_fps(seed_points, self.num_proposal))
        elif sample_mod == 'random':
            B, S = seed_points.shape[:2]
            idx = torch.randint(0, S, (B, self.num_proposal), device=seed_points.device, dtype=torch.int32)
            agg_in = dict(points_xyz=vote_points, features=vote_features, indices=idx)
        else:
            agg_in = dict(points_xyz=seed_points, features=seed_features, target_xyz=vote_points)
        aggregated_points, features, aggregated_indices = self._k_aggregate(**agg_in)
        results['aggregated_points'] = aggregated_points
        results['aggregated_features'] = features
        results['aggregated_indices'] = aggregated_indices

        rows, B, P = self.conv_pred.forward_rows(features)
        rows = rows.view(B, P, -1)
        n_cls = self.num_classes + 2
        results['obj_scores'] = rows[..., :2]
        results['sem_scores'] = rows[..., 2:n_cls]
        results = self.side2box(aggregated_points, rows[..., n_cls:], results)

        center, size, heading, results = self.jitter_bbox_preds(results, dataset_name, jitter_noise)
        results = self.grid_conv(center.detach(), size.detach(), heading.detach(), results)
        iou = results['iou_scores'].sigmoid()
        results['iou_scores_jitter'], results['iou_scores'] = iou[:, P:], iou[:, :P]
        side = results['side_scores'].sigmoid().permute(1, 3, 0, 2)
        results['side_scores_jitter'], results['side_scores'] = side[:, P:], side[:, :P]
        return results

    # ---- test-time decoding -----------------------------------------------------------------------
    def _k_points_in_boxes_count(self, points, boxes):
        """points (B, N, 3) depth frame, boxes (B, P, 7) bottom-centred -> (B, P) points per box
        (DepthInstance3DBoxes.points_in_boxes: depth -> LiDAR flip + points_in_boxes_batch)."""
        from .points_in_boxes import points_in_boxes_batch
        pl = torch.stack([points[..., 1], -points[..., 0], points[..., 2]], dim=-1).contiguous()
        bl = torch.stack([boxes[..., 1], -boxes[..., 0], boxes[..., 2], boxes[..., 4], boxes[..., 3],
                          boxes[..., 5], boxes[..., 6]], dim=-1).contiguous()
        return points_in_boxes_batch(pl, bl).sum(dim=1)

    def _k_aligned_nms(self, boxes, scores, classes, thresh, counts):
        from .box3d_nms import aligned_3d_nms_batched
        return aligned_3d_nms_batched(boxes, scores, classes, thresh, counts)

    def get_bboxes(self, points, bbox_preds, input_metas=None, rescale=False, use_nms=True,
                   use_iou_for_nms=True):
        """NesieHead.get_bboxes / multiclass_nms_single (nesie_head.py:681-788) for all scenes at once:
        objectness x IoU score, axis-aligned extent of every (rotated) box, non-empty test (> 5 points
        inside), class-aware aligned_3d_nms over the non-empty boxes (ONE launch for the batch), score
        threshold, per-class expansion.  Returns [(boxes (k, 7) bottom-centred, scores (k,), labels (k,))]
        per scene; the only host sync is the final split into variable-length lists."""
        cfg = self.test_cfg
        obj_scores = F.softmax(bbox_preds['obj_scores'], dim=-1)[..., -1]
        sem_scores = F.softmax(bbox_preds['sem_scores'], dim=-1)
        bbox3d = bbox_preds['bbox_preds']
        if use_iou_for_nms:
            indx = bbox_preds['sem_scores'].argmax(dim=-1, keepdim=True)
            obj_scores = obj_scores * torch.gather(bbox_preds['iou_scores'], 2, indx).squeeze(-1)
        if not use_nms:
            return bbox3d
        B, P = obj_scores.shape
        # box_type_3d(bbox, origin=(0.5, 0.5, 0.5)): gravity centre -> bottom centre
        boxes = bbox3d.clone()
        boxes[..., 2] = boxes[..., 2] + boxes[..., 5] * (0.0 - 0.5)
        nonempty = self._k_points_in_boxes_count(points[..., :3].contiguous(), boxes) > 5
        # corners (depth_box3d.py:51-89) -> axis-aligned min / max
        dims = boxes[..., 3:6]
        norm = boxes.new_tensor([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1],
                                 [1, 1, 1], [1, 1, 0]]) - boxes.new_tensor([0.5, 0.5, 0])
        corners = dims.unsqueeze(2) * norm.view(1, 1, 8, 3)
        sin, cos = torch.sin(boxes[..., 6]).unsqueeze(-1), torch.cos(boxes[..., 6]).unsqueeze(-1)
        cx = corners[..., 0] * cos + corners[..., 1] * sin
        cy = corners[..., 0] * (-sin) + corners[..., 1] * cos
        corners = torch.stack([cx, cy, corners[..., 2]], dim=-1) + boxes[..., None, :3]
        minmax = torch.cat([corners.min(dim=2)[0], corners.max(dim=2)[0]], dim=-1)
        classes = torch.argmax(sem_scores, -1)
        # aligned_3d_nms over the non-empty boxes only: compact them to the front (stable)
        order = torch.argsort((~nonempty).to(torch.int8), dim=1, stable=True)
        g = lambda t: torch.gather(t, 1, order if t.dim() == 2 else order.unsqueeze(-1).expand(-1, -1, t.shape[-1]))  # noqa: E731
        keep, keep_cnt = self._k_aligned_nms(g(minmax), g(obj_scores), g(classes), cfg['nms_thr'],
                                             nonempty.sum(dim=1))
        valid = torch.arange(P, device=keep.device).unsqueeze(0) < keep_cnt.unsqueeze(1)
        picked = torch.zeros((B, P + 1), dtype=torch.bool, device=keep.device)
        src = torch.where(valid, torch.gather(order, 1, keep.clamp(min=0)), torch.full_like(keep, P))
        picked.scatter_(1, src, True)
        selected = picked[:, :P] & (obj_scores > cfg['score_thr'])
        results = []
        for b in range(B):
            sel = selected[b]
            bs, sc, cl = boxes[b][sel], obj_scores[b][sel], classes[b][sel]
            if cfg.get('per_class_proposal', False):
                C = sem_scores.shape[-1]
                sem = sem_scores[b][sel]
                results.append((bs.repeat(C, 1), (sc.unsqueeze(0) * sem.t()).reshape(-1),
                                torch.arange(C, device=cl.device).repeat_interleave(bs.shape[0])))
            else:
                results.append((bs, sc, cl))
        return results

    # ---- targets --------------------------------------------------------------------------------
    def get_targets_padded(self, points, boxes, labels, valid, bbox_preds, at_seeds=True):
        """points (B, N, >=3); boxes (B, G, 7) bottom-centred, valid rows first; labels (B, G);
        valid (B, G) bool.  Returns a dict with the reference's target tensors; the vote targets
        cover the seeds only when at_seeds (all the loss reads), every point otherwise."""
        agg = bbox_preds['aggregated_points']
        boxes = boxes * valid.unsqueeze(-1).to(boxes.dtype)
        n_valid = valid.sum(dim=1).to(torch.int32)
        n_slots = n_valid.clamp(min=1)                 # an empty scene is one fake zero box
        vote_targets, vote_masks = self._k_vote_targets(
            points, boxes, n_valid, bbox_preds['seed_indices'].long() if at_seeds else None)
        gc = T.gravity_center(boxes)
        assignment, _ = self._k_chamfer_assign(agg, gc, n_slots, (True, False))
        a3 = assignment.unsqueeze(-1)
        d1 = ((agg - torch.gather(gc, 1, a3.expand(-1, -1, 3))) ** 2).sum(-1)
        dist = torch.sqrt(d1.detach() + 1e-6)
        pos = dist < self.train_cfg['pos_distance_thr']
        objectness_targets = pos.long()
        objectness_masks = (pos | (dist > self.train_cfg['neg_distance_thr'])).float()
        mask_targets = torch.gather(labels, 1, assignment)
        bbox_targets = torch.cat([torch.gather(gc, 1, a3.expand(-1, -1, 3)),
                                  torch.gather(boxes[..., 3:], 1, a3.expand(-1, -1, 4))], dim=-1)
        valid_f = valid.float()
        return dict(
            vote_targets=vote_targets, vote_target_masks=vote_masks, center_targets=gc,
            bbox_targets=bbox_targets, mask_targets=mask_targets, valid_gt_masks=valid.long(),
            objectness_targets=objectness_targets,
            objectness_weights=objectness_masks / (torch.sum(objectness_masks) + 1e-6),
            box_loss_weights=objectness_targets.float() / (torch.sum(objectness_targets).float() + 1e-6),
            valid_gt_weights=valid_f / (torch.sum(valid_f) + 1e-6), assignment=assignment,
            max_gt_num=n_slots.max().expand(n_slots.shape[0]).contiguous(), at_seeds=at_seeds)

    def get_targets(self, points, gt_bboxes_3d, gt_labels_3d, pts_semantic_mask=None,
                    pts_instance_mask=None, bbox_preds=None):
        """The reference's list interface and return tuple (vote targets for every point; tensors
        padded to the batch's largest box count)."""
        dev = bbox_preds['aggregated_points'].device
        pts = torch.stack(list(points)) if isinstance(points, (list, tuple)) else points
        boxes, labels, valid = T.pad_gt(gt_bboxes_3d, gt_labels_3d, dev)
        t = self.get_targets_padded(pts.to(dev), boxes, labels, valid, bbox_preds, at_seeds=False)
        return (t['vote_targets'], t['vote_target_masks'], t['center_targets'],
                list(t['bbox_targets']), t['mask_targets'], t['valid_gt_masks'],
                t['objectness_targets'], t['objectness_weights'], t['box_loss_weights'],
                t['valid_gt_weights'], t['assignment'])

    # ---- loss pieces ----------------------------------------------------------------------------
    def _center_loss(self, bbox_preds, t):
        cfg = self.loss_cfg['center']
        pc, gc = bbox_preds['bbox_preds'][..., :3], t['center_targets']
        i1, i2 = self._k_chamfer_assign(pc, gc, t['max_gt_num'])
        src = ((pc - torch.gather(gc, 1, i1.unsqueeze(-1).expand(-1, -1, 3))) ** 2).sum(-1)
        dst = ((torch.gather(pc, 1, i2.unsqueeze(-1).expand(-1, -1, 3)) - gc) ** 2).sum(-1)
        return torch.sum(src * t['box_loss_weights']) * cfg.get('loss_src_weight', 1.0) + \
            torch.sum(dst * t['valid_gt_weights']) * cfg.get('loss_dst_weight', 1.0)

    def _semantic_loss(self, bbox_preds, t):
        ce = F.cross_entropy(bbox_preds['sem_scores'].transpose(2, 1), t['mask_targets'], reduction='none')
        return self.loss_cfg['semantic'].get('loss_weight', 1.0) * (ce * t['box_loss_weights']).sum()

    def _sigma_terms(self, bbox_preds, t, surface_weight):
        """Surface loss with per-side uncertainty (fused kernel) -> (surface_loss, sigma (rows, 6))."""
        C = bbox_preds['side_scores'].shape[-1]
        return self._k_side_loss(bbox_preds['surface_pred'].reshape(-1, 6), t['bbox_targets'].reshape(-1, 7),
                                 bbox_preds['side_scores'].reshape(-1, 6, C),
                                 bbox_preds['sem_scores'].reshape(-1, C), surface_weight)

    def _iou_loss(self, bbox_preds, t, sigma, iou_weight):
        lw = self.loss_cfg['iou'].get('loss_weight', 1.0)
        iou = cal_iou_3d(bbox_preds['bbox_preds'], t['bbox_targets'], self._k_sort_vertices).reshape(-1)
        per_row = lw * ((1 - iou) * iou_weight)
        sigma_mean = sigma.mean(dim=-1)
        if self.uncertainty == 'saqe':
            return (torch.exp(-sigma_mean.detach()) * per_row).sum(), iou
        return (torch.exp(-sigma_mean) * per_row + self.alpha * sigma_mean * iou_weight).sum(), iou

    # ---- supervised loss ------------------------------------------------------------------------
    def _loss_streams(self):
        return int(os.environ.get("NESIE_LOSS_STREAMS", "5"))

    def loss_padded(self, bbox_preds, points, boxes, labels, valid, ret_target=False):
        """All eight loss terms.  After the (sequential) target assignment the terms are independent
        chains of small kernels: on CUDA they run as forked branches (branches.py)."""
        t = self.get_targets_padded(points, boxes, labels, valid, bbox_preds)
        C = self.num_classes
        box_w = t['box_loss_weights'].reshape(-1)
        surface_weight = box_w.unsqueeze(-1).repeat(1, 6)
        label_cls = t['mask_targets'].reshape(-1)
        qcfg = self.loss_cfg['iou_pred']

        def vote():
            return dict(vote_loss=self.vote_module.get_loss(
                bbox_preds['seed_points'], bbox_preds['vote_points'], bbox_preds['seed_indices'],
                t['vote_target_masks'], t['vote_targets'], at_seeds=t['at_seeds']))

        def objectness():
            ocfg = self.loss_cfg['objectness']
            ce = F.cross_entropy(bbox_preds['obj_scores'].transpose(2, 1), t['objectness_targets'],
                                 weight=self._obj_class_weight, reduction='none')
            return dict(objectness_loss=ocfg.get('loss_weight', 1.0) * (ce * t['objectness_weights']).sum())

        def center_semantic():
            return dict(center_loss=self._center_loss(bbox_preds, t),
                        semantic_loss=self._semantic_loss(bbox_preds, t))

        def surface_iou():
            # surface loss with per-side uncertainty (fused kernel), IoU loss with its mean, and the
            # IoU-score regression (quality focal loss) that needs the same IoU
            surface_loss, sigma = self._sigma_terms(bbox_preds, t, surface_weight)
            iou_loss, iou = self._iou_loss(bbox_preds, t, sigma, box_w)
            qfl = quality_focal_loss(bbox_preds['iou_scores'].reshape(-1, C), label_cls, iou.detach(), box_w,
                                     qcfg.get('beta', 2.0), qcfg.get('loss_weight', 1.0))
            return dict(surface_loss=surface_loss, iou_loss=iou_loss, _qfl=qfl)

        def jitter_side():
            label_iou_jitter = cal_iou_3d(bbox_preds['jitter_bbox_preds'].detach(), t['bbox_targets'],
                                          self._k_sort_vertices).reshape(-1)
            qfl = quality_focal_loss(bbox_preds['iou_scores_jitter'].reshape(-1, C), label_cls,
                                     label_iou_jitter, box_w, qcfg.get('beta', 2.0),
                                     qcfg.get('loss_weight', 1.0))
            # side-score regression: label = min(4 |surface - target|, 1) (SidePredLoss)
            side_pred = torch.gather(bbox_preds['side_scores'].reshape(-1, 6, C), 2,
                                     label_cls.view(-1, 1, 1).expand(-1, 6, 1)).squeeze(-1)
            label_side = (4.0 * (bbox_preds['surface_pred'].reshape(-1, 6)
                                 - bbox2surface(t['bbox_targets'].reshape(-1, 7))).abs()).detach().clamp(max=1.0)
            side_loss = self.loss_cfg['side'].get('loss_weight', 1.0) * \
                (F.mse_loss(side_pred, label_side, reduction='none') * surface_weight).sum()
            return dict(_qfl_jitter=qfl, side_loss=side_loss)

        shared = [v for v in list(bbox_preds.values()) + list(t.values()) if torch.is_tensor(v)]
        shared += [box_w, surface_weight, label_cls]
        parts = run_branches([surface_iou, jitter_side, center_semantic, vote, objectness],
                             bbox_preds['aggregated_points'], shared, self._loss_streams())
        r = {k: v for d in parts for k, v in d.items()}
        losses = dict(vote_loss=r['vote_loss'], objectness_loss=r['objectness_loss'],
                      semantic_loss=r['semantic_loss'], center_loss=r['center_loss'],
                      surface_loss=r['surface_loss'], iou_loss=r['iou_loss'],
                      iou_pred_loss=r['_qfl'] + r['_qfl_jitter'], side_loss=r['side_loss'])
        if ret_target:
            losses['targets'] = t['bbox_targets']
        return losses

    def loss(self, bbox_preds, points, gt_bboxes_3d, gt_labels_3d, pts_semantic_mask=None,
             pts_instance_mask=None, img_metas=None, gt_bboxes_ignore=None, ret_target=False):
        dev = bbox_preds['aggregated_points'].device
        pts = torch.stack(list(points)) if isinstance(points, (list, tuple)) else points
        boxes, labels, valid = T.pad_gt(gt_bboxes_3d, gt_labels_3d, dev)
        return self.loss_padded(bbox_preds, pts.to(dev), boxes, labels, valid, ret_target)

    # ---- unsupervised loss ----------------------------------------------------------------------
    def unsup_loss_padded(self, bbox_preds, points, boxes, labels, valid, quality, un_label_weight=2.0):
        """Pseudo-label losses (nesie_head.py:414-509).  quality (B, G, 6): per-side quality of every
        pseudo box, zero rows where invalid."""
        t = self.get_targets_padded(points, boxes, labels, valid, bbox_preds)
        a = t['assignment']
        quality = quality * valid.unsqueeze(-1).to(quality.dtype)
        q_side = torch.gather(quality, 1, a.unsqueeze(-1).expand(-1, -1, 6))
        q_mean = q_side.mean(dim=-1)
        box_w = t['box_loss_weights']
        surface_weight = box_w.reshape(-1).unsqueeze(-1).repeat(1, 6) * q_side.reshape(-1, 6)
        iou_weight = (box_w * q_mean).reshape(-1)

        def surface_iou():
            surface_loss, sigma = self._sigma_terms(bbox_preds, t, surface_weight)
            return surface_loss, self._iou_loss(bbox_preds, t, sigma, iou_weight)[0]

        shared = [v for v in list(bbox_preds.values()) + list(t.values()) if torch.is_tensor(v)]
        shared += [surface_weight, iou_weight]
        (surface_loss, iou_loss), center_loss, semantic_loss = run_branches(
            [surface_iou, lambda: self._center_loss(bbox_preds, t), lambda: self._semantic_loss(bbox_preds, t)],
            bbox_preds['aggregated_points'], shared, self._loss_streams())
        return dict(unsup_semantic_loss=un_label_weight * semantic_loss,
                    unsup_center_loss=un_label_weight * center_loss,
                    unsup_iou_loss=un_label_weight * iou_loss,
                    unsup_surface_loss=un_label_weight * surface_loss)

    def unsup_loss(self, bbox_preds, points, pseudo_boxes, pseudo_label, img_metas=None,
                   pseudo_quality_score=None):
        dev = bbox_preds['aggregated_points'].device
        pts = torch.stack(list(points)) if isinstance(points, (list, tuple)) else points
        boxes, labels, valid, quality = T.pad_gt(pseudo_boxes, pseudo_label, dev,
                                                 extra=pseudo_quality_score)
        return self.unsup_loss_padded(bbox_preds, pts.to(dev), boxes, labels, valid, quality)
