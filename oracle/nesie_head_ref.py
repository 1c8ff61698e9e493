"""CPU twin of nesie_b200.nesie_head.NesieHead: identical torch glue, every kernel hook replaced by a
CPU restatement that follows the reference line by line.  TEST INFRASTRUCTURE ONLY (parity tests,
bench.py's cpu_baseline / `--impl reference` legs).  Pinned by tests/test_head_cpu.py against
tests/golden/head_golden.npz, which the reference's own NesieHead source produced
(tests/golden/make_golden_head.py)."""
import torch

from nesie_b200.nesie_head import NesieHead

from . import cpu
from . import modules as om
from . import restate


def vote_targets_ref(points, boxes, n_valid, seed_indices=None):
    """NesieHead.get_targets_single, vote part (nesie_head.py:618-654), one scene and one GT box at a
    time like the reference; points_in_boxes through DepthInstance3DBoxes.points_in_boxes
    (depth_box3d.py:251-277: depth -> LiDAR flip of points and boxes, then points_in_boxes_batch)."""
    B, N, _ = points.shape
    out_t, out_m = [], []
    for b in range(B):
        pts = points[b]
        n = int(n_valid[b])
        vote_targets = pts.new_zeros([N, 9])
        vote_target_masks = pts.new_zeros([N], dtype=torch.long)
        vote_target_idx = pts.new_zeros([N], dtype=torch.long)
        if n:
            bx = boxes[b, :n]
            pl = pts[:, [1, 0, 2]].clone()
            pl[:, 1] *= -1
            rt = bx.new_tensor([[0, 1, 0], [-1, 0, 0], [0, 0, 1]])
            bl = torch.cat([bx[:, :3] @ rt.t(), bx[:, [4, 3, 5]], bx[:, 6:]], dim=-1)
            inside = cpu.points_in_boxes_batch(pl[None], bl[None])[0]
            gc = bx[:, :3].clone()
            gc[:, 2] = bx[:, 2] + bx[:, 5] * 0.5
            for i in range(n):
                indices = torch.nonzero(inside[:, i], as_tuple=False).squeeze(-1)
                vote_target_masks[indices] = 1
                tmp = vote_targets[indices]
                votes = gc[i].unsqueeze(0) - pts[indices, :3]
                for j in range(3):
                    col = torch.nonzero(vote_target_idx[indices] == j, as_tuple=False).squeeze(-1)
                    tmp[col, 3 * j:3 * j + 3] = votes[col]
                    if j == 0:
                        tmp[col] = votes[col].repeat(1, 3)
                vote_targets[indices] = tmp
                vote_target_idx[indices] = torch.clamp(vote_target_idx[indices] + 1, max=2)
        if seed_indices is not None:
            vote_targets, vote_target_masks = vote_targets[seed_indices[b]], vote_target_masks[seed_indices[b]]
        out_t.append(vote_targets)
        out_m.append(vote_target_masks)
    return torch.stack(out_t), torch.stack(out_m)


def chamfer_assign_ref(src, dst, n_valid=None, want=(True, True)):
    """chamfer_distance (losses/chamfer_distance.py:49-56, l2) per scene over the first n_valid[b]
    destination points for the source -> destination argmin."""
    src, dst = src.detach(), dst.detach()
    B, N, _ = src.shape
    M = dst.shape[1]
    i1 = torch.zeros((B, N), dtype=torch.long, device=src.device)
    i2 = torch.zeros((B, M), dtype=torch.long, device=src.device)
    for b in range(B):
        s = src[b].unsqueeze(1).repeat(1, M, 1)
        d = dst[b].unsqueeze(0).repeat(N, 1, 1)
        dist = torch.nn.functional.mse_loss(s, d, reduction='none').sum(-1)
        nv = M if n_valid is None else int(n_valid[b])
        i1[b] = torch.min(dist[:, :nv], dim=1)[1]
        i2[b] = torch.min(dist, dim=0)[1]
    return (i1 if want[0] else None), (i2 if want[1] else None)


class NesieHeadOracle(NesieHead):

    @staticmethod
    def _k_side_pooling_cls():
        from .side_pooling_ref import SidePoolingOracle
        return SidePoolingOracle

    def _k_aggregate(self, points_xyz, features=None, indices=None, target_xyz=None):
        return om.sa_forward(self.vote_aggregation, points_xyz, features, indices, target_xyz)

    def _k_fps(self, xyz, n):
        return cpu.furthest_point_sample(xyz, n)

    def _k_vote_targets(self, points, boxes, n_valid, seed_indices):
        return vote_targets_ref(points, boxes, n_valid, seed_indices)

    def _k_chamfer_assign(self, src, dst, n_valid, want=(True, True)):
        return chamfer_assign_ref(src, dst, n_valid, want)

    _k_sort_vertices = staticmethod(cpu.sort_vertices)

    def _k_points_in_boxes_count(self, points, boxes):
        pl = torch.stack([points[..., 1], -points[..., 0], points[..., 2]], dim=-1)
        bl = torch.stack([boxes[..., 1], -boxes[..., 0], boxes[..., 2], boxes[..., 4], boxes[..., 3],
                          boxes[..., 5], boxes[..., 6]], dim=-1)
        return cpu.points_in_boxes_batch(pl, bl).to(points.device).sum(dim=1)

    def _k_aligned_nms(self, boxes, scores, classes, thresh, counts):
        """aligned_3d_nms per scene (core/post_processing/box3d_nms.py:129-176), numpy restatement."""
        B, P = scores.shape
        keep = torch.full((B, P), -1, dtype=torch.long)
        cnt = torch.zeros((B,), dtype=torch.int32)
        for b in range(B):
            n = int(counts[b])
            if n:
                k = restate.aligned_3d_nms(boxes[b, :n].detach().cpu().numpy(), scores[b, :n].detach().cpu().numpy(),
                                           classes[b, :n].cpu().numpy(), thresh)
                keep[b, :len(k)] = torch.from_numpy(k)
                cnt[b] = len(k)
        return keep.to(scores.device), cnt.to(scores.device)

    def _k_side_loss(self, surface_pred, box_targets, side_scores, sem_scores, weight):
        lw = self.loss_cfg['surface'].get('loss_weight', 1.0)
        if self.uncertainty == 'saqe':
            return restate.side_uncertainty_loss(surface_pred, box_targets, side_scores.detach(),
                                                 sem_scores, weight, lw, 0.0)
        return restate.side_uncertainty_loss(surface_pred, box_targets, side_scores, sem_scores,
                                             weight, lw, self.alpha)
