"""Generates tests/golden/head_golden.npz from the REFERENCE's own NesieHead source
(mmdet3d/models/dense_heads/nesie_head.py: forward :211-275, loss :277-412, unsup_loss :414-509,
get_targets :511-591, get_targets_single :593-679) lifted unmodified by ref_lift.py, together with
the reference's VoteModule.get_loss, ChamferDistance, SurfaceLoss, SidePredLoss, IoU3DLoss
(cal_iou_3d of ops/rotated_iou), GeneralQualityFocalLoss and DepthInstance3DBoxes.

    python tests/golden/make_golden_head.py        (build container only: reads /root/reference)

Stored: the inputs, every loss term, every target tensor and the gradients of the summed loss with
respect to the differentiable predictions ("loss*" cases); for the "fwd" case the module is
re-created from the seed (same construction order / state_dict layout as the reference) and the
outputs of NesieHead.forward in train and eval mode are stored with a parameter checksum."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_lift  # noqa: E402

OUT = os.path.join(HERE, "head_golden.npz")
SEED = 20261019
NUM_CLASSES, REG_MAX = 18, 32


def make_scene(g, n_pts, n_gt, P, S):
    """One scene: points (n_pts, 4), bottom-centred yaw-0 GT boxes (n_gt, 7), labels."""
    pts = torch.rand(n_pts, 4, generator=g) * torch.tensor([6.0, 6.0, 2.5, 1.0]) - torch.tensor([3.0, 3.0, 0.0, 0.0])
    if n_gt:
        ctr = torch.rand(n_gt, 3, generator=g) * torch.tensor([5.0, 5.0, 1.0]) - torch.tensor([2.5, 2.5, 0.0])
        size = torch.rand(n_gt, 3, generator=g) * 1.6 + 0.4
        boxes = torch.cat([ctr, size, torch.zeros(n_gt, 1)], dim=1)
        labels = torch.randint(0, NUM_CLASSES, (n_gt,), generator=g)
        # overlapping boxes so that some points lie in >= 4 boxes (gt_per_seed bookkeeping)
        if n_gt >= 5:
            boxes[1:5, :3] = boxes[0, :3] + (torch.rand(4, 3, generator=g) - 0.5) * 0.3
        # a quarter of the points are drawn inside boxes
        k = n_pts // 4
        which = torch.randint(0, n_gt, (k,), generator=g)
        u = torch.rand(k, 3, generator=g) - 0.5
        pts[:k, :3] = boxes[which, :3] + u * boxes[which, 3:6] + torch.tensor([0, 0, 0.5]) * boxes[which, 5:6]
    else:
        boxes = torch.zeros(0, 7)
        labels = torch.zeros(0, dtype=torch.long)
    return pts, boxes, labels


def make_preds(g, points, gts, P, S):
    B = len(points)
    d = {}
    seed_idx = torch.stack([torch.randperm(p.shape[0], generator=g)[:S] for p in points])
    d["seed_indices"] = seed_idx
    d["seed_points"] = torch.stack([p[i, :3] for p, i in zip(points, seed_idx)])
    d["vote_points"] = d["seed_points"] + torch.randn(B, S, 3, generator=g) * 0.3
    agg = torch.rand(B, P, 3, generator=g) * torch.tensor([6.0, 6.0, 2.0]) - torch.tensor([3.0, 3.0, 0.0])
    box_c = torch.zeros(B, P, 3)
    box_s = torch.rand(B, P, 3, generator=g) * 1.5 + 0.3
    for b, gt in enumerate(gts):
        n = gt.shape[0]
        for p in range(P):
            if n and p % 2 == 0:      # half of the proposals sit near a GT centre
                k = p // 2 % n
                c = gt[k, :3] + torch.tensor([0.0, 0.0, 0.5]) * gt[k, 5]
                agg[b, p] = c + torch.randn(3, generator=g) * (0.1 if p % 4 == 0 else 0.35)
                box_s[b, p] = gt[k, 3:6] * (1 + 0.3 * (torch.rand(3, generator=g) - 0.5))
        box_c[b] = agg[b] + torch.randn(P, 3, generator=g) * 0.08
    yaw = torch.randn(B, P, 1, generator=g) * 0.4
    yaw[:, ::5] = 0.0
    d["aggregated_points"] = agg
    d["bbox_preds"] = torch.cat([box_c, box_s, yaw], dim=-1)
    d["jitter_bbox_preds"] = torch.cat([box_c + box_s * torch.randn(B, P, 3, generator=g) * 0.3,
                                        (box_s * (1 + torch.randn(B, P, 3, generator=g) * 0.3)).clamp(min=1e-8),
                                        yaw], dim=-1)
    d["surface_pred"] = torch.cat([box_c - 0.5 * box_s, box_c + 0.5 * box_s], dim=-1) \
        + torch.randn(B, P, 6, generator=g) * 0.02
    d["surface_scale"] = torch.tensor([3.0, 3.0, 2.5, 3.0, 3.0, 2.5]).expand(B, P, 6).contiguous()
    d["bbox_probs"] = torch.softmax(torch.randn(B, 6, REG_MAX + 1, P, generator=g), dim=2)
    d["obj_scores"] = torch.randn(B, P, 2, generator=g)
    d["sem_scores"] = torch.randn(B, P, NUM_CLASSES, generator=g)
    d["iou_scores"] = torch.rand(B, P, NUM_CLASSES, generator=g) * 0.98 + 0.01
    d["iou_scores_jitter"] = torch.rand(B, P, NUM_CLASSES, generator=g) * 0.98 + 0.01
    d["side_scores"] = torch.rand(B, P, 6, NUM_CLASSES, generator=g)
    return d


GRAD_KEYS = ["vote_points", "bbox_preds", "surface_pred", "obj_scores", "sem_scores", "iou_scores",
             "iou_scores_jitter", "side_scores"]
TARGET_NAMES = ["vote_targets", "vote_target_masks", "center_targets", "bbox_targets", "mask_targets",
                "valid_gt_masks", "objectness_targets", "objectness_weights", "box_loss_weights",
                "valid_gt_weights", "assignment"]


def run_loss_case(ns, head, out, tag, gt_counts, n_pts, P, S, seed):
    g = torch.Generator().manual_seed(seed)
    Boxes = ns["DepthInstance3DBoxes"]
    scenes = [make_scene(g, n_pts, n, P, S) for n in gt_counts]
    points = [s[0] for s in scenes]
    gts = [s[1] for s in scenes]
    labels = [s[2] for s in scenes]
    preds = make_preds(g, points, gts, P, S)
    out[f"{tag}_shape"] = np.array([len(scenes), n_pts, P, S], dtype=np.int64)
    out[f"{tag}_gt_counts"] = np.array(gt_counts, dtype=np.int64)
    out[f"{tag}_points"] = torch.stack(points).numpy()
    out[f"{tag}_gt_boxes"] = torch.cat(gts).numpy() if sum(gt_counts) else np.zeros((0, 7), np.float32)
    out[f"{tag}_gt_labels"] = torch.cat(labels).numpy()
    for k, v in preds.items():
        out[f"{tag}_in_{k}"] = v.numpy()
    # ---- supervised loss -------------------------------------------------------------------
    p = {k: (v.clone().requires_grad_(True) if k in GRAD_KEYS else v.clone()) for k, v in preds.items()}
    gt_list = [Boxes(b.clone(), box_dim=7, with_yaw=True) for b in gts]
    lab_list = [l.clone() for l in labels]
    losses = head.loss(p, [x.clone() for x in points], gt_list, lab_list)
    total = sum(losses.values())
    total.backward()
    for k, v in losses.items():
        out[f"{tag}_sup_{k}"] = v.detach().numpy()
    for k in GRAD_KEYS:
        out[f"{tag}_sup_grad_{k}"] = (p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])).numpy()
    gt_list = [Boxes(b.clone(), box_dim=7, with_yaw=True) for b in gts]
    lab_list = [l.clone() for l in labels]
    targets = head.get_targets([x.clone() for x in points], gt_list, lab_list, bbox_preds=preds)
    for name, t in zip(TARGET_NAMES, targets):
        if name == "bbox_targets":
            t = torch.cat(t, dim=0)
        out[f"{tag}_tgt_{name}"] = t.numpy()
    # ---- unsupervised loss: pseudo boxes / labels / per-side quality -------------------------
    # (the pseudo boxes are bottom-centred like GT; empty scenes carry float labels as
    # get_pseudo_labels builds them, votenet_nesie.py:289-292)
    pl_boxes, pl_labels, pl_quality = [], [], []
    for b, gt in enumerate(gts):
        n = gt.shape[0]
        keep = n if b % 2 == 0 else max(n - 2, 0)
        jit = gt[:keep].clone()
        jit[:, :6] += torch.randn(keep, 6, generator=g) * 0.05
        pl_boxes.append(jit)
        pl_labels.append(labels[b][:keep].clone() if keep else torch.ones((0,)))
        pl_quality.append(torch.rand(keep, 6, generator=g))
    out[f"{tag}_pl_counts"] = np.array([x.shape[0] for x in pl_boxes], dtype=np.int64)
    out[f"{tag}_pl_boxes"] = torch.cat(pl_boxes).numpy()
    out[f"{tag}_pl_labels"] = torch.cat([x.float() for x in pl_labels]).numpy()
    out[f"{tag}_pl_quality"] = torch.cat(pl_quality).numpy()
    p = {k: (v.clone().requires_grad_(True) if k in GRAD_KEYS else v.clone()) for k, v in preds.items()}
    losses = head.unsup_loss(p, [x.clone() for x in points],
                             [Boxes(b.clone(), box_dim=7, with_yaw=True) for b in pl_boxes],
                             [l.clone() for l in pl_labels], None, [q.clone() for q in pl_quality])
    total = sum(losses.values())
    total.backward()
    for k, v in losses.items():
        out[f"{tag}_unsup_{k}"] = v.detach().numpy()
    for k in GRAD_KEYS:
        out[f"{tag}_unsup_grad_{k}"] = (p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])).numpy()


def small_head_cfg(C, P, mean_path):
    cfg = dict(ref_lift.NESIE_HEAD_CFG)
    cfg["vote_module_cfg"] = dict(cfg["vote_module_cfg"], in_channels=C, conv_channels=(C, C))
    cfg["vote_aggregation_cfg"] = dict(cfg["vote_aggregation_cfg"], num_point=P, radius=0.6,
                                       num_sample=8, mlp_channels=[C, 32, 32, 32])
    cfg["pred_layer_cfg"] = dict(in_channels=32, shared_conv_channels=(32, 32), bias=True)
    cfg["grid_conv_cfg"] = dict(num_class=NUM_CLASSES, num_heading_bin=1, num_size_cluster=NUM_CLASSES,
                                mean_size_arr_path=mean_path, num_proposal=P, sampling="seed_fps",
                                query_feats="seed", seed_feat_dim=C)
    return cfg


def main():
    ref_lift.patch_cuda_noop()
    ns = ref_lift.base_namespace()
    tmp = os.path.join(tempfile.mkdtemp(), "mean.npz")
    np.savez(tmp, np.ones((NUM_CLASSES, 3), dtype=np.float32))
    out = {"seed": np.int64(SEED)}

    # ---- loss / target cases: the head's parameters are not involved -------------------------
    torch.manual_seed(SEED)
    head = ns["NesieHead"](**small_head_cfg(16, 8, tmp))
    run_loss_case(ns, head, out, "lossA", gt_counts=[6, 0, 3], n_pts=240, P=16, S=32, seed=SEED + 1)
    run_loss_case(ns, head, out, "lossB", gt_counts=[1, 9], n_pts=300, P=24, S=40, seed=SEED + 2)

    # ---- forward case ------------------------------------------------------------------------
    C, P, S, B = 16, 8, 48, 2
    torch.manual_seed(SEED + 10)
    head = ns["NesieHead"](**small_head_cfg(C, P, tmp))
    g = torch.Generator().manual_seed(SEED + 11)
    seed_points = torch.rand(B, S, 3, generator=g) * torch.tensor([3.0, 3.0, 1.5])
    seed_feats = torch.randn(B, C, S, generator=g)
    seed_idx = torch.stack([torch.randperm(400, generator=g)[:S] for _ in range(B)])
    out["fwd_shape"] = np.array([B, S, C, P], dtype=np.int64)
    out["fwd_seed_points"] = seed_points.numpy()
    out["fwd_seed_features"] = seed_feats.numpy()
    out["fwd_seed_indices"] = seed_idx.numpy()
    out["fwd_param_abs_sum"] = np.float64(sum(p.detach().double().abs().sum() for p in head.parameters()))
    out["fwd_keys"] = np.array(sorted(head.state_dict().keys()))
    keys = ["vote_points", "vote_features", "aggregated_points", "aggregated_features",
            "aggregated_indices", "obj_scores", "sem_scores", "surface_pred", "surface_scale",
            "bbox_preds", "bbox_probs", "jitter_bbox_preds", "iou_scores", "iou_scores_jitter",
            "side_scores", "side_scores_jitter"]
    for mode in ("train", "eval"):
        head.train(mode == "train")
        feat = dict(fp_xyz=[seed_points.clone()], fp_features=[seed_feats.clone()],
                    fp_indices=[seed_idx.clone()])
        torch.manual_seed(SEED + 12)     # the jitter noise: two torch.randn(B, P, 3) draws (:185-186)
        with torch.no_grad():
            res = head(feat, "vote", "ScanNet")
        for k in keys:
            out[f"fwd_{mode}_{k}"] = res[k].numpy()
    # ---- test-time decoding: get_bboxes / multiclass_nms_single (:681-788) ------------------------
    import types
    ref_lift.lift(os.path.join(ref_lift.REF, "core/post_processing/box3d_nms.py"), ns, names={"aligned_3d_nms"})
    g = torch.Generator().manual_seed(SEED + 20)
    B, N, P = 2, 3000, 48
    pts = torch.rand(B, N, 3, generator=g) * torch.tensor([4.0, 4.0, 2.0])
    nclu = 5
    clu = torch.rand(B, nclu, 3, generator=g) * torch.tensor([3.0, 3.0, 1.0]) + 0.5
    which = torch.randint(0, nclu, (B, P), generator=g)
    ctr = torch.gather(clu, 1, which.unsqueeze(-1).expand(-1, -1, 3)) + torch.randn(B, P, 3, generator=g) * 0.1
    size = torch.rand(B, P, 3, generator=g) * 0.8 + 0.3
    size[:, ::7] = 0.02                                     # boxes that hold no points
    yaw = torch.randn(B, P, 1, generator=g) * 0.3
    preds = dict(bbox_preds=torch.cat([ctr, size, yaw], -1),
                 obj_scores=torch.randn(B, P, 2, generator=g) * 2,
                 sem_scores=torch.randn(B, P, NUM_CLASSES, generator=g) * 2,
                 iou_scores=torch.rand(B, P, NUM_CLASSES, generator=g))
    for per_class in (True, False):
        head.test_cfg = types.SimpleNamespace(nms_thr=0.25, score_thr=0.05, per_class_proposal=per_class)
        metas = [dict(box_type_3d=ns["DepthInstance3DBoxes"]) for _ in range(B)]
        res = head.get_bboxes(pts, {k: v.clone() for k, v in preds.items()}, metas)
        tag = f"dec{int(per_class)}"
        out[f"{tag}_counts"] = np.array([r[1].shape[0] for r in res], dtype=np.int64)
        out[f"{tag}_boxes"] = torch.cat([r[0].tensor for r in res]).numpy()
        out[f"{tag}_scores"] = torch.cat([r[1] for r in res]).numpy()
        out[f"{tag}_labels"] = torch.cat([r[2] for r in res]).numpy()
        print(tag, "boxes per scene:", out[f"{tag}_counts"])
    out["dec_points"] = pts.numpy()
    for k, v in preds.items():
        out[f"dec_in_{k}"] = v.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
