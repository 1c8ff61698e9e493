"""CPU twins of nesie_b200.detectors.VoteNet / VoteNetNesie: identical step structure, every kernel
replaced by the CPU oracle (C restatement of the reference kernels, torch-CPU MLPs, loop-form
pseudo-label filter, per-tensor EMA as SimiTeacherHook does it).  TEST INFRASTRUCTURE ONLY: parity
tests and bench.py's cpu_baseline / `--impl reference` legs."""
import torch

from nesie_b200 import detectors as D
from nesie_b200.pointnet2_sa_ssg import PointNet2SASSG

from . import modules as om
from . import restate
from .nesie_head_ref import NesieHeadOracle


class BackboneRef(PointNet2SASSG):
    def forward(self, points, fps_indices=None, after_level=None):
        return om.backbone_forward(self, points)


class VoteNetRef(D.VoteNet):
    head_cls = NesieHeadOracle
    backbone_cls = BackboneRef


class TeacherEMARef:
    """core/utils/simi_teacher_hook.py:39-92 on CPU tensors, one tensor at a time."""

    def __init__(self, model, momentum=0.001, interval=1, warm_up=10):
        self.momentum, self.interval, self.warm_up = momentum ** interval, interval, warm_up
        self.params = [p for _, p in model.named_parameters(recurse=True)]
        self.ema = [p.data.clone() for p in self.params]

    def after_train_iter(self, curr_step):
        if curr_step % self.interval != 0:
            return
        for e, p in zip(self.ema, self.params):
            restate.ema_update(e, p.data, self.momentum, self.warm_up, curr_step)

    def swap(self):
        for e, p in zip(self.ema, self.params):
            temp = p.data.clone()
            p.data.copy_(e)
            e.copy_(temp)


class VoteNetNesieRef(D.VoteNetNesie):
    head_cls = NesieHeadOracle
    backbone_cls = BackboneRef

    def init_teacher(self):
        self.teacher = TeacherEMARef(self, **self.ema_cfg)
        return self.teacher

    def teacher_pseudo_labels(self, points_t, aug_t=None, aug_s=None, **kw):
        with torch.no_grad():
            self.teacher.swap()
            preds_t = self.predict(points_t, **kw)
            labels, boxes, quality = restate.get_pseudo_labels(
                preds_t, self.ulb_list, self.ulb_flag, self.n_lb, self.n_ulb,
                num_classes=self.bbox_head.num_classes,
                thresh_warmup=self.train_cfg.get('thresh_warmup', True),
                use_cbl=self.train_cfg.get('use_cbl', True), quality_poly=self.quality_poly)
            B, G = len(boxes), D.MAX_NUM_OBJ
            pb = torch.zeros(B, G, 7)
            pl = torch.zeros(B, G, dtype=torch.long)
            pv = torch.zeros(B, G, dtype=torch.bool)
            pq = torch.zeros(B, G, 6)
            for i in range(B):
                n = boxes[i].shape[0]
                if n:
                    pb[i, :n], pl[i, :n], pv[i, :n], pq[i, :n] = boxes[i], labels[i].long(), True, quality[i]
            dev = points_t.device        # (the reference transforms the boxes on the CPU, :313-316)
            pb, pl, pv, pq = pb.to(dev), pl.to(dev), pv.to(dev), pq.to(dev)
            pb = D.transformation_bbox_preds(pb, aug_t, aug_s) * pv.unsqueeze(-1)
            self.teacher.swap()
        return pb, pl, pv, pq
