"""Shared loaders for the NesieHead golden cases (tests/golden/head_golden.npz)."""
import os
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
NUM_CLASSES = 18
GRAD_KEYS = ["vote_points", "bbox_preds", "surface_pred", "obj_scores", "sem_scores", "iou_scores",
             "iou_scores_jitter", "side_scores"]
TARGET_NAMES = ["vote_targets", "vote_target_masks", "center_targets", "bbox_targets", "mask_targets",
                "valid_gt_masks", "objectness_targets", "objectness_weights", "box_loss_weights",
                "valid_gt_weights", "assignment"]
LOSS_KEYS = ["vote_loss", "objectness_loss", "semantic_loss", "center_loss", "surface_loss", "iou_loss",
             "iou_pred_loss", "side_loss"]
UNSUP_KEYS = ["unsup_semantic_loss", "unsup_center_loss", "unsup_iou_loss", "unsup_surface_loss"]

# configs/Nesie/nesie-votenet-scannet-train-010.py:17-93 (reference NesieHead arguments)
HEAD_CFG = dict(
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


def load_golden():
    return np.load(os.path.join(HERE, "golden", "head_golden.npz"))


def mean_size_file():
    path = os.path.join(tempfile.mkdtemp(), "mean.npz")
    np.savez(path, np.ones((NUM_CLASSES, 3), dtype=np.float32))
    return path


def make_head(cls, C, P, **kw):
    """The reduced head of make_golden_head.py::small_head_cfg."""
    cfg = dict(HEAD_CFG)
    cfg["vote_module_cfg"] = dict(cfg["vote_module_cfg"], in_channels=C, conv_channels=(C, C))
    cfg["vote_aggregation_cfg"] = dict(cfg["vote_aggregation_cfg"], num_point=P, radius=0.6,
                                       num_sample=8, mlp_channels=[C, 32, 32, 32])
    cfg["pred_layer_cfg"] = dict(in_channels=32, shared_conv_channels=(32, 32), bias=True)
    cfg["grid_conv_cfg"] = dict(num_class=NUM_CLASSES, num_heading_bin=1, num_size_cluster=NUM_CLASSES,
                                mean_size_arr_path=mean_size_file(), num_proposal=P,
                                sampling="seed_fps", query_feats="seed", seed_feat_dim=C)
    cfg.update(kw)
    return cls(**cfg)


def loss_inputs(G, tag, device, grad=False):
    """-> (bbox_preds dict, points (B, N, 4), gt box list, gt label list, (pl boxes, labels, quality))."""
    preds = {}
    for k in G.files:
        if k.startswith(f"{tag}_in_"):
            t = torch.from_numpy(G[k]).to(device)
            name = k[len(tag) + 4:]
            if grad and name in GRAD_KEYS:
                t.requires_grad_(True)
            preds[name] = t
    points = torch.from_numpy(G[f"{tag}_points"]).to(device)

    def split(arr, counts):
        out, o = [], 0
        for c in counts:
            out.append(torch.from_numpy(arr[o:o + c]).to(device))
            o += c
        return out
    counts = [int(c) for c in G[f"{tag}_gt_counts"]]
    boxes = split(G[f"{tag}_gt_boxes"], counts)
    labels = split(G[f"{tag}_gt_labels"], counts)
    pc = [int(c) for c in G[f"{tag}_pl_counts"]]
    pl = (split(G[f"{tag}_pl_boxes"], pc), [t.long() for t in split(G[f"{tag}_pl_labels"], pc)],
          split(G[f"{tag}_pl_quality"], pc))
    return preds, points, boxes, labels, pl


def rel_err(a, b):
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def decode_inputs(G, device):
    preds = {k[len("dec_in_"):]: torch.from_numpy(G[k]).to(device) for k in G.files if k.startswith("dec_in_")}
    return torch.from_numpy(G["dec_points"]).to(device), preds


def check_decoded(G, per_class, results):
    """get_bboxes results against the reference's: same boxes in the same order, scores to 1e-6."""
    tag = f"dec{int(per_class)}"
    counts = [int(c) for c in G[f"{tag}_counts"]]
    assert [r[1].shape[0] for r in results] == counts
    o = 0
    for (b, s, l), n in zip(results, counts):
        assert torch.equal(l.cpu().long(), torch.from_numpy(G[f"{tag}_labels"][o:o + n]).long())
        assert torch.allclose(b.cpu(), torch.from_numpy(G[f"{tag}_boxes"][o:o + n]), rtol=1e-6, atol=1e-6)
        assert torch.allclose(s.cpu(), torch.from_numpy(G[f"{tag}_scores"][o:o + n]), rtol=2e-6, atol=1e-7)
        o += n
