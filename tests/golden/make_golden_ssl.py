"""Generates tests/golden/ssl_golden.npz from the REFERENCE's own mean-teacher source:

  * VoteNetNesie.get_pseudo_labels (models/detectors/votenet_nesie.py:129-299, with
    lhs_3d_faster_samecls / get_3d_box / roty / flip_axis_to_camera :733-821), run as a whole with a
    stub `self` (ulb_list, ulb_flag, lb_map, ulb_map, CLASSES, train_cfg);
  * VoteNetNesie.ulb_update (:301-308), choose_sup_item / choose_unsup_item (:46-67),
    transformation_bbox_preds / untransformation / transformation (:310-324, 596-634) on the
    reference's DepthInstance3DBoxes;
  * SimiTeacherHook (core/utils/simi_teacher_hook.py:39-92): EMA buffers, update, swap.

    python tests/golden/make_golden_ssl.py        (build container only: reads /root/reference)

One property of the environment is fixed on purpose: get_pseudo_labels picks its 64 candidates with
`torch.argsort(pos_obj * iou * final_mask, descending=True)` (:206), whose key is exactly 0 for every
masked-out proposal; the order among those ties is unspecified for torch's default (unstable) sort
and differs between the CPU and CUDA backends, yet it decides which masked-out boxes take part in
the NMS.  torch.argsort is made STABLE while the reference code runs, so the vectors record the
reference's algorithm under the stable tie order (the one nesie_b200 uses).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_lift  # noqa: E402

OUT = os.path.join(HERE, "ssl_golden.npz")
SEED = 20261020
C = 18


def make_teacher_preds(g, B, P, easy):
    """Teacher predictions; `easy` scenes have many proposals that pass every threshold."""
    nclu = 6
    clu_c = torch.rand(B, nclu, 3, generator=g) * torch.tensor([5.0, 5.0, 1.5]) - torch.tensor([2.5, 2.5, 0.0])
    which = torch.randint(0, nclu, (B, P), generator=g)
    center = torch.gather(clu_c, 1, which.unsqueeze(-1).expand(-1, -1, 3)) + torch.randn(B, P, 3, generator=g) * 0.15
    size = torch.rand(B, P, 3, generator=g) * 1.0 + 0.5
    heading = torch.randn(B, P, 1, generator=g) * 0.2
    sem = torch.randn(B, P, C, generator=g)
    # the class follows the cluster for most proposals, with a confident logit
    cls = (which * 3) % C
    boost = torch.where(torch.rand(B, P, generator=g) < (0.9 if easy else 0.5),
                        torch.full((B, P), 2.5), torch.zeros(B, P))
    sem.scatter_add_(2, cls.unsqueeze(-1), boost.unsqueeze(-1))
    obj = torch.randn(B, P, 2, generator=g) * (1.0 if easy else 2.0)
    obj[..., 1] += 3.0 if easy else 1.0
    iou = torch.rand(B, P, C, generator=g) * (0.7 if easy else 0.6) + (0.3 if easy else 0.05)
    side = torch.rand(B, P, 6, C, generator=g)
    vote = torch.randn(B, P, 3, generator=g)
    return dict(bbox_preds=torch.cat([center, size, heading], -1), sem_scores=sem, obj_scores=obj,
                iou_scores=iou, side_scores=side, vote_points=vote)


def main():
    ref_lift.patch_cuda_noop()
    ns = ref_lift.base_namespace()
    ns["SingleStage3DDetector"] = object
    ns["bbox3d2result"] = ns["merge_aug_bboxes_3d"] = None
    ref_lift.lift("models/detectors/votenet_nesie.py", ns)
    Det = ns["VoteNetNesie"]
    Boxes = ns["DepthInstance3DBoxes"]
    out = {"seed": np.int64(SEED)}

    # ---- get_pseudo_labels + ulb_update -------------------------------------------------------
    for case, (B, P, easy, warm, n_ulb, n_lb) in enumerate([(4, 160, True, True, 40, 10),
                                                            (3, 96, False, True, 25, 5),
                                                            (2, 128, True, False, 12, 3)]):
        g = torch.Generator().manual_seed(SEED + case)
        preds = make_teacher_preds(g, B, P, easy)
        ulb_list = torch.randint(0, 4, (n_ulb, C), generator=g).float()
        ulb_list[:, ::5] = 0
        ulb_flag = (torch.rand(n_ulb, generator=g) < 0.6).float()
        stub = types.SimpleNamespace(
            ulb_list=ulb_list.clone(), ulb_flag=ulb_flag.clone(), lb_map=list(range(n_lb)),
            ulb_map=[1000 + i for i in range(n_ulb)], CLASSES=tuple(f"c{i}" for i in range(C)),
            train_cfg=types.SimpleNamespace(thresh_warmup=warm, use_cbl=True))
        tag = f"pl{case}"
        out[f"{tag}_cfg"] = np.array([B, P, int(warm), n_ulb, n_lb], dtype=np.int64)
        for k, v in preds.items():
            out[f"{tag}_in_{k}"] = v.numpy().copy()
        out[f"{tag}_ulb_list"] = ulb_list.numpy()
        out[f"{tag}_ulb_flag"] = ulb_flag.numpy()
        p = {k: v.clone() for k, v in preds.items()}
        unstable_argsort = torch.argsort
        torch.argsort = lambda *a, **k: unstable_argsort(*a, **{**k, "stable": True})
        try:
            labels, boxes, quality = Det.get_pseudo_labels(stub, p, "ScanNet")
        finally:
            torch.argsort = unstable_argsort
        out[f"{tag}_counts"] = np.array([b.shape[0] for b in boxes], dtype=np.int64)
        out[f"{tag}_labels"] = torch.cat([l.float() for l in labels]).numpy()
        out[f"{tag}_boxes"] = torch.cat(boxes).numpy()
        out[f"{tag}_quality"] = torch.cat(quality).numpy()
        out[f"{tag}_bbox_preds_after"] = p["bbox_preds"].numpy()     # shifted in place (:152)
        # ulb_update with the scenes' positions in the unlabeled table
        pos = torch.randperm(n_ulb, generator=g)[:B].tolist()
        metas = [dict(sample_idx=stub.ulb_map[q]) for q in pos]
        Det.ulb_update(stub, labels, metas)
        out[f"{tag}_ulb_pos"] = np.array(pos, dtype=np.int64)
        out[f"{tag}_ulb_list_after"] = stub.ulb_list.numpy()
        out[f"{tag}_ulb_flag_after"] = stub.ulb_flag.numpy()
        print(tag, "pseudo boxes per scene:", out[f"{tag}_counts"])

    # ---- choose_sup_item / choose_unsup_item --------------------------------------------------
    g = torch.Generator().manual_seed(SEED + 20)
    use_label = [True, False, True, False, False]
    preds = dict(a=torch.randn(5, 4, 3, generator=g), b=torch.randn(5, 2, generator=g))
    inputs = ([torch.full((2,), float(i)) for i in range(5)], [torch.full((1,), 10.0 + i) for i in range(5)], None)
    sup_p, sup_in = Det.choose_sup_item(None, preds, inputs, use_label)
    unsup_p, unsup_in = Det.choose_unsup_item(None, preds, inputs[:2], use_label)
    out["choose_use_label"] = np.array(use_label)
    out["choose_a"], out["choose_b"] = preds["a"].numpy(), preds["b"].numpy()
    out["choose_sup_a"], out["choose_unsup_a"] = sup_p["a"].numpy(), unsup_p["a"].numpy()
    out["choose_sup_in0"] = torch.stack(sup_in[0]).numpy()
    out["choose_unsup_in1"] = torch.stack(unsup_in[1]).numpy()
    assert sup_in[2] is None

    # ---- transformation_bbox_preds: teacher frame -> original -> student frame ------------------
    g = torch.Generator().manual_seed(SEED + 30)
    nb = [5, 0, 3, 7]
    boxes = [torch.cat([torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g) + 0.3,
                        torch.randn(n, 1, generator=g) * 0.3], -1) for n in nb]

    def meta(flow):
        ang = float(torch.rand(1, generator=g) * 0.17 - 0.085)
        c, s = np.cos(ang), np.sin(ang)
        return dict(transformation_3d_flow=flow,
                    pcd_rotation=torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=torch.float32).T,
                    pcd_scale_factor=float(torch.rand(1, generator=g) * 0.2 + 0.9),
                    pcd_trans=(torch.randn(3, generator=g) * 0.1).numpy())
    flows_t = [["HF", "VF", "R", "S", "T"], ["R", "S", "T"], ["VF", "R", "S", "T"], ["HF", "R", "S", "T"]]
    flows_s = [["VF", "R", "S", "T"], ["HF", "VF", "R", "S", "T"], ["R", "S", "T"], ["HF", "VF", "R", "S", "T"]]
    metas_t, metas_s = [meta(f) for f in flows_t], [meta(f) for f in flows_s]
    res = Det.transformation_bbox_preds(types.SimpleNamespace(
        untransformation=lambda b, m: Det.untransformation(None, b, m),
        transformation=lambda b, m: Det.transformation(None, b, m)), [b.clone() for b in boxes], metas_t, metas_s)
    out["tf_counts"] = np.array(nb, dtype=np.int64)
    out["tf_boxes"] = torch.cat(boxes).numpy()
    out["tf_result"] = torch.cat([r.tensor for r in res]).numpy()
    for side, metas in (("t", metas_t), ("s", metas_s)):
        out[f"tf_{side}_hf"] = np.array(["HF" in m["transformation_3d_flow"] for m in metas])
        out[f"tf_{side}_vf"] = np.array(["VF" in m["transformation_3d_flow"] for m in metas])
        out[f"tf_{side}_rot"] = np.stack([m["pcd_rotation"].numpy() for m in metas])
        out[f"tf_{side}_scale"] = np.array([m["pcd_scale_factor"] for m in metas], dtype=np.float32)
        out[f"tf_{side}_trans"] = np.stack([m["pcd_trans"] for m in metas]).astype(np.float32)

    # ---- SimiTeacherHook ----------------------------------------------------------------------
    ref_lift.lift("core/utils/simi_teacher_hook.py", ns)
    torch.manual_seed(SEED + 40)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    hook = ns["SimiTeacherHook"](momentum=0.001, interval=1, warm_up=10)
    hook.hooks_before_run(model)
    out["ema_buffer_names"] = np.array(sorted(n for n, _ in model.named_buffers() if n.startswith("ema_")))
    g = torch.Generator().manual_seed(SEED + 41)
    steps = []
    for it in range(4):
        with torch.no_grad():
            for p_ in model.parameters():
                p_.add_(torch.randn(p_.shape, generator=g) * 0.1)
        hook.hooks_after_train_iter(it)
        steps.append(torch.cat([model.state_dict()[hook.param_ema_buffer[n]].reshape(-1)
                                for n, _ in model.named_parameters()]).clone())
    out["ema_params_final"] = torch.cat([p_.detach().reshape(-1) for p_ in model.parameters()]).numpy()
    out["ema_after_each_step"] = torch.stack(steps).numpy()
    hook._swap_ema_parameters()
    out["ema_swapped_params"] = torch.cat([p_.detach().reshape(-1) for p_ in model.parameters()]).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
