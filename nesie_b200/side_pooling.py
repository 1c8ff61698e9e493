"""SidePooling quality head (reference: mmdet3d/models/dense_heads/side_pooling_module.py) on this
repo's kernels -- SURVEY.md 8f-1, the step right after vote aggregation.

Same constructor arguments, sub-module names (`mlps_before.{i}.first_conv.*`, `mlps_head.{i}.*`:
checkpoints load unchanged), inputs and outputs as the reference class.  What changes is how the
forward is evaluated on CUDA tensors:

  grid features (`grid_features`, :183-243)   three_nn kernel for the 3 nearest seeds of every grid
      point + `nesie_interp_rows`: inverse-distance interpolation written straight as row-major
      GEMM rows [relative xyz | C features] (the reference gathers with a python index_select loop
      over the batch and materialises a (B, K*G*3, C) tensor);
  MiniPointNet (:343-370)                     1x1 Conv2d stacks as fp32-parity row GEMMs on tcgen05
      with the training BatchNorm fused in (mlp_rows.py), max over the grid points of a box as a
      row-group max;
  mlps_head (:54-80)                          Conv1d / BatchNorm1d stacks as row GEMMs.

center / size / heading arrive detached (nesie_head.py:264) and the seeds are detached inside
(:83-86), so gradients only reach this module's own parameters.  There is no CPU path: the hooks
`_grid_rows`, `_mini_pointnet`, `_head` are what oracle/side_pooling_ref.py overrides with a CPU
restatement of the reference arithmetic."""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import bn_rows
from . import gather_linear
from . import group_max
from . import mlp_rows
from . import pool_rows
from .interpolate import three_nn, three_nn_grid
from .linear_rows import linear_rows


def rot_gpu(t):
    """(...,) angles -> (..., 3, 3) rotations about the upright axis (side_pooling_module.py:326-340)."""
    out = torch.zeros(tuple(t.shape) + (3, 3), dtype=t.dtype, device=t.device)
    c, s = torch.cos(t), torch.sin(t)
    out[..., 0, 0] = c
    out[..., 0, 1] = s
    out[..., 1, 0] = -s
    out[..., 1, 1] = c
    out[..., 2, 2] = 1
    return out


class _GridSource:
    """Input rows of one MiniPointNet, [grid point - box centre | features interpolated from the 3
    nearest seeds | 0 pad], in factored form: the dense (B * n, ld) tensor is only built on demand."""

    def __init__(self, xyz_count, table, idx, weight, head, ld):
        self.m = xyz_count                 # seeds per scene
        self.table = table                 # (B, m, C) point-major seed features
        self.idx, self.weight, self.head = idx, weight, head      # (B, n, 3) each
        self.ld = ld

    @property
    def shape(self):
        return (self.idx.shape[0] * self.idx.shape[1], self.ld)

    def tensors(self):
        return [self.table, self.idx, self.weight, self.head]

    def rows(self):
        B, n = self.idx.shape[:2]
        C = self.table.shape[2]
        with torch.no_grad():
            rows = torch.empty((B * n, self.ld), dtype=torch.float32, device=self.table.device)
            with torch.cuda.device(self.table.device):
                _lib.call("nesie_interp_rows", B, C, self.m, n, _lib.ptr(self.table), _lib.ptr(self.idx),
                          _lib.ptr(self.weight), _lib.ptr(self.head), _lib.ptr(rows), self.ld, _lib.stream())
        return rows


def _gather_linear_enabled():
    return os.environ.get("NESIE_GATHER_LINEAR", "1") != "0"


class MiniPointNet(nn.Module):
    """Parameter container with the reference's layout (:343-358); evaluated by SidePooling."""

    def __init__(self, channels, feature_dim, hide_dim=256):
        super().__init__()
        self.first_conv = nn.Sequential(
            nn.Conv2d(channels, hide_dim, 1, bias=False), nn.BatchNorm2d(hide_dim),
            nn.ReLU(inplace=True), nn.Conv2d(hide_dim, hide_dim // 2, 1))
        self.second_conv = nn.Sequential(
            nn.Conv2d(hide_dim, hide_dim, 1, bias=False), nn.BatchNorm2d(hide_dim),
            nn.ReLU(inplace=True), nn.Conv2d(hide_dim, feature_dim, 1))

    def forward(self, points):
        """(B, C, K, G) -> (B, feature_dim, K), the reference's tensor formulation (:360-370)."""
        feature = self.first_conv(points)
        feature_global = torch.max(feature, dim=-1, keepdim=True).values
        feature = torch.cat([feature_global.expand(-1, -1, -1, feature.shape[-1]), feature], dim=1)
        feature = self.second_conv(feature)
        return torch.max(feature, dim=-1).values


def _pad_cols(x, mult=4):
    pad = (-x.shape[1]) % mult
    return F.pad(x, (0, pad)) if pad else x


def _conv_bn_relu_conv(rows, conv_a, bn, conv_b, add_bias=True):
    """rows (R, Ka) -> conv_b(relu(bn(conv_a(rows)))) as (R, Nb); conv_a has no bias.
    add_bias=False leaves conv_b's bias to the caller (it is folded into the group-max kernel)."""
    wa = conv_a.weight.flatten(1)
    if rows.shape[1] != wa.shape[1]:           # zero-padded input columns
        wa = F.pad(wa, (0, rows.shape[1] - wa.shape[1]))
    wb = conv_b.weight.flatten(1)
    if bn.training and mlp_rows.supported(rows, [(wa, bn)]) and wb.shape[0] % 4 == 0 and wb.shape[0] <= 256:
        y1, parts = mlp_rows._LinearStats.apply(rows, wa)
        rm, rv = mlp_rows._bn_buffers(bn)
        y2, _ = mlp_rows._BNReLULinear.apply(y1, parts, bn.weight, bn.bias, rm, rv, bn.eps,
                                             bn.momentum, wb)
    elif (not bn.training and not torch.is_grad_enabled() and bn.track_running_stats and
          mlp_rows.supported_eval(rows, wa, wb)):
        # inference: the BatchNorm is the affine map y * s + t of its running statistics, applied
        # with the ReLU in the second GEMM's operand prologue (no elementwise sweep at all)
        scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        y1, _ = mlp_rows._gemm_fused(rows, wa, None, None, False)
        y2, _ = mlp_rows._gemm_fused(y1, wb, scale.contiguous(), shift.contiguous(), False)
    else:
        y1 = linear_rows(rows, wa)
        if bn.training and bn_rows.supported(y1, bn):
            a = bn_rows.bn_relu_rows(y1, bn)
        else:
            if bn.training and bn.track_running_stats:
                bn_rows.count_batch(bn)
            a = F.relu(F.batch_norm(y1, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                                    bn.training, bn.momentum, bn.eps))
        y2 = linear_rows(a, wb)
    return y2 + conv_b.bias if (add_bias and conv_b.bias is not None) else y2


class SidePooling(nn.Module):
    """Side / IoU quality estimation from features interpolated on a grid over every proposal box."""

    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr_path,
                 num_proposal, sampling, seed_feat_dim=256, query_feats="seed",
                 iou_class_depend=True):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = None if mean_size_arr_path is None else \
            np.load(mean_size_arr_path)["arr_0"]
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.query_feats = query_feats
        self.iou_class_depend = iou_class_depend
        self.reg_topk = 4
        self.grid_size = g = 4
        self.left_mask = [i // g * g * g + i % g for i in range(g * g)]
        self.right_mask = [i // g * g * g + i % g + g * (g - 1) for i in range(g * g)]
        self.iou_size = num_class if iou_class_depend else 1
        # device copies of the face masks (not in the state_dict): indexing with the python lists
        # would upload an index tensor on every call, which CUDA-graph capture forbids
        self.register_buffer("_left_idx", torch.tensor(self.left_mask), persistent=False)
        self.register_buffer("_right_idx", torch.tensor(self.right_mask), persistent=False)

        before, head = [], []
        for _ in range(6):   # six sides, then the whole box (same creation order as the reference)
            before.append(MiniPointNet(seed_feat_dim + 3, 128))
            head.append(nn.Sequential(
                nn.Conv1d(128 + 33 + 4 + 1, 128, 1), nn.BatchNorm1d(128), nn.ReLU(),
                nn.Conv1d(128, 128, 1), nn.BatchNorm1d(128), nn.ReLU(),
                nn.Conv1d(128, self.iou_size, 1)))
        before.append(MiniPointNet(seed_feat_dim + 3, 128))
        head.append(nn.Sequential(
            nn.Conv1d(128, 128, 1), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, 128, 1), nn.BatchNorm1d(128), nn.ReLU(),
            nn.Conv1d(128, self.iou_size, 1)))
        self.mlps_before = nn.ModuleList(before)
        self.mlps_head = nn.ModuleList(head)

    # ---- geometry (small elementwise work; reference :83-181) --------------------------------
    def extract_features(self, end_points):
        return (end_points["seed_points"].detach().contiguous(),
                end_points["seed_features"].detach().contiguous())

    def generate_grid(self, size):
        """size (B, K, 3) -> (B, K, g^3, 3) box-frame grid, x slowest / z fastest."""
        g = self.grid_size
        step = torch.linspace(-1, 1, g, device=size.device, dtype=size.dtype)
        gx = step.view(g, 1, 1).expand(g, g, g).reshape(1, 1, -1)
        gy = step.view(1, g, 1).expand(g, g, g).reshape(1, 1, -1)
        gz = step.view(1, 1, g).expand(g, g, g).reshape(1, 1, -1)
        return torch.stack([gx * size[:, :, 0:1] / 2, gy * size[:, :, 1:2] / 2,
                            gz * size[:, :, 2:3] / 2], dim=-1)

    def _to_world(self, grid, center, heading):
        B, K = center.shape[:2]
        rot = rot_gpu(heading).view(-1, 3, 3)
        out = torch.bmm(grid.reshape(B * K, -1, 3), rot.transpose(1, 2)).view(B, K, -1, 3)
        return out + center.unsqueeze(2)

    def grid_for_side(self, whole_grid, center, heading):
        """-> (B, K, 6 * g^2, 3): front, back, top, down, left, right faces of the grid."""
        g = self.grid_size
        faces = [whole_grid[:, :, 0:g * g], whole_grid[:, :, -g * g:], whole_grid[:, :, g - 1::g],
                 whole_grid[:, :, ::g], whole_grid.index_select(2, self._left_idx),
                 whole_grid.index_select(2, self._right_idx)]
        return self._to_world(torch.cat(faces, dim=-2), center, heading)

    def grid_for_bbox(self, whole_grid, center, heading):
        return self._to_world(whole_grid, center, heading)

    def dist_feature(self, end_points, prefix=''):
        """(B, 6, 33, K/2) side distributions -> (6, B, 38, K): probs, top-4, variance (:245-264)."""
        prob = end_points[f"{prefix}bbox_probs"].detach()
        stat = torch.cat([prob, prob.topk(self.reg_topk, dim=2)[0], prob.var(dim=2, keepdim=True)],
                         dim=2)
        return stat.permute(1, 0, 2, 3).repeat(1, 1, 1, 2)

    # ---- hot path hooks ----------------------------------------------------------------------
    def _grid_rows(self, origin_xyz, origin_features, grid, center, sides=1, nn_cache=None):
        """grid (B, T, 3) world points, T = K * sides * G ordered (box, side, grid point) -> per side the
        rows (B*K*G, ld) = [grid - centre | features interpolated from the 3 nearest seeds with
        normalised inverse-distance weights | 0 pad] as a _GridSource (factored: seed table + 3
        neighbours + weights + relative position); one source when sides == 1, else a list.
        Each side gets its own contiguous block (the MiniPointNet of a side reads only its rows):
        the 3-NN search runs once over all grid points."""
        _lib.need_cuda(origin_xyz, origin_features, grid, center)
        B, T = grid.shape[:2]
        K = center.shape[1]
        G = T // (K * sides)
        C = origin_features.shape[1]
        with torch.no_grad():
            if origin_xyz.shape[1] >= 64 and os.environ.get("NESIE_THREE_NN_GRID", "0") == "1":
                # exact 3-NN through a grid over the seeds, binned once per forward (nn_cache).  Off by
                # default: measured (tools/nn_probe.py, 8 x 81920 grid points, 1024 seeds) 304 us against
                # 443 us brute force when every grid point lies within ~0.2 m of a seed, 473 us at ~0.6 m
                # and 1.7 ms at ~2 m (the shells a thread walks diverge and grow); the boxes of an
                # untrained or early-training head put most grid points in the last regime.
                ws = nn_cache.get("ws") if nn_cache is not None else None
                dist, idx, ws = three_nn_grid(grid, origin_xyz, ws)
                if nn_cache is not None:
                    nn_cache["ws"] = ws
            else:
                dist, idx = three_nn(grid, origin_xyz)
            weight = 1.0 / (dist + 1e-8)
            weight = weight / weight.sum(dim=2, keepdim=True)
            head = grid.view(B, K, sides * G, 3) - center.unsqueeze(2)
            table = origin_features.transpose(1, 2).contiguous()          # (B, N, C) point-major
            ld = -(-(3 + C) // 4) * 4
            out = []
            for i in range(sides):
                pick = lambda t: t.view(B, K, sides, G, 3)[:, :, i].reshape(B, K * G, 3).contiguous()  # noqa: E731
                out.append(_GridSource(origin_xyz.shape[1], table, pick(idx), pick(weight), pick(head), ld))
        return out[0] if sides == 1 else out

    def _side_rows(self, origin_xyz, origin_features, side_grid, center, nn_cache=None):
        """Rows of the six face grids, one contiguous block per side."""
        return self._grid_rows(origin_xyz, origin_features, side_grid, center, sides=6, nn_cache=nn_cache)

    def _mini_pointnet(self, mpn, rows, G):
        """rows (R, ld) -- a tensor or a _GridSource -- with every G consecutive rows one box ->
        (R / G, feature_dim)."""
        fused = self._mini_pointnet_pooled(mpn, rows, G)
        if fused is not None:
            return fused
        if isinstance(rows, _GridSource):
            rows = rows.rows()
        feat = _conv_bn_relu_conv(rows, mpn.first_conv[0], mpn.first_conv[1], mpn.first_conv[3],
                                  add_bias=False)
        # [max over the box's grid points, broadcast | feature] with the conv bias folded in
        feat = group_max.group_max_concat_rows(feat, mpn.first_conv[3].bias, G)
        feat = _conv_bn_relu_conv(feat, mpn.second_conv[0], mpn.second_conv[1], mpn.second_conv[3],
                                  add_bias=False)
        return group_max.group_max_rows(feat, mpn.second_conv[3].bias, G)

    @staticmethod
    def _mini_pointnet_pooled(mpn, rows, G):
        """Training path with both maxima taken in the GEMM epilogues (pool_rows.py); None when the
        shapes / modes are not covered (the caller then runs the step-by-step formulation)."""
        (ca, bna, _, cb), (cc, bnc, _, cd) = mpn.first_conv, mpn.second_conv
        wa, wb = ca.weight.flatten(1), cb.weight.flatten(1)
        wc, wd = cc.weight.flatten(1), cd.weight.flatten(1)
        C = wb.shape[0]
        R = rows.shape[0]
        if not (bna.training and bnc.training and pool_rows.enabled() and pool_rows.pool_unit(G) > 0 and
                R % G == 0 and (G & (G - 1)) == 0 and
                bna.affine and bna.momentum is not None and bnc.affine and bnc.momentum is not None and
                wa.shape[0] % 4 == 0 and wa.shape[0] <= 256 and wb.shape[1] == wa.shape[0] and
                C % 4 == 0 and C <= 256 and wc.shape[1] == 2 * C and
                cb.bias is not None and cd.bias is not None and
                wc.shape[0] % 4 == 0 and wc.shape[0] <= 256 and wd.shape[1] == wc.shape[0] and
                wd.shape[0] % 4 == 0 and wd.shape[0] <= 256):
            return None
        if (isinstance(rows, _GridSource) and _gather_linear_enabled() and
                gather_linear.supported(wa.shape[0]) and wa.shape[1] == 3 + rows.table.shape[2]):
            # conv_a commuted with the interpolation: its GEMM runs over the seeds, the grid rows gather
            # the transformed seeds (gather_linear.cu); the seed features carry no gradient (:83-95)
            src = rows
            B, m, Cs = src.table.shape
            seeds = linear_rows(src.table.view(B * m, Cs), wa[:, 3:].contiguous())
            y1, parts1 = gather_linear.gather_linear(seeds.view(B, m, -1), src.idx, src.weight, src.head,
                                                     wa[:, :3], True)
        else:
            if isinstance(rows, _GridSource):
                rows = rows.rows()
            if rows.shape[1] != wa.shape[1]:           # zero-padded input columns
                wa = F.pad(wa, (0, rows.shape[1] - wa.shape[1]))
            if not mlp_rows.supported(rows, [(wa, bna)]):
                return None
            y1, parts1 = mlp_rows._LinearStats.apply(rows, wa)
        ya, gmax, arg = pool_rows.bn_relu_linear_max(y1, parts1, bna, wb, cb.bias, G, True)
        yc, partsc = pool_rows.concat_global_linear(ya, gmax, arg, cb.bias, wc, G, zero_mean_grad=True)
        return pool_rows.bn_relu_linear_max(yc, partsc, bnc, wd, cd.bias, G, False)

    def _head(self, seq, x):
        """Conv1d / BatchNorm1d / ReLU stack on x (B, C, K) -> (B, C_out, K), as row GEMMs."""
        B, C, K = x.shape
        r = x.transpose(1, 2).reshape(B * K, C)
        mods = list(seq)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv1d):
                w = m.weight.flatten(1)
                rp = _pad_cols(r)
                if rp.shape[1] != w.shape[1]:
                    w = F.pad(w, (0, rp.shape[1] - w.shape[1]))
                r = linear_rows(rp.contiguous(), w)
                if m.bias is not None:
                    r = pool_rows.add_bias_rows(r, m.bias)
                i += 1
            elif isinstance(m, nn.BatchNorm1d):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                if relu and m.training and bn_rows.supported(r, m):
                    r = bn_rows.bn_relu_rows(r, m)
                else:
                    if m.training and m.track_running_stats:
                        bn_rows.count_batch(m)
                    r = F.batch_norm(r, m.running_mean, m.running_var, m.weight, m.bias, m.training,
                                     m.momentum, m.eps)
                    if relu:
                        r = F.relu(r)
                i += 2 if relu else 1
            else:
                r = m(r)
                i += 1
        return r.view(B, K, -1).transpose(1, 2)

    # ---- forward (reference :266-323) --------------------------------------------------------
    def forward(self, center, size, heading, end_points, prefix=''):
        """center / size (B, K, 3), heading (B, K); end_points holds seed_points (B, N, 3),
        seed_features (B, C, N) and {prefix}bbox_probs (B, 6, 33, K/2).  Adds
        {prefix}side_scores (6, B, iou_size, K) and {prefix}iou_scores (B, K, iou_size)."""
        B, K = size.shape[:2]
        g2, g3 = self.grid_size ** 2, self.grid_size ** 3
        origin_xyz, origin_features = self.extract_features(end_points)
        whole_grid = self.generate_grid(size)
        side_grid = self.grid_for_side(whole_grid, center, heading).reshape(B, -1, 3).contiguous()
        bbox_grid = self.grid_for_bbox(whole_grid, center, heading).reshape(B, -1, 3).contiguous()
        nn_cache = {}      # the seeds are binned once for both grids
        side_rows = self._side_rows(origin_xyz, origin_features, side_grid, center, nn_cache=nn_cache)
        bbox_rows = self._grid_rows(origin_xyz, origin_features, bbox_grid, center, nn_cache=nn_cache)
        dist_feature = self.dist_feature(end_points, prefix)

        def branch(i):
            if i < 6:
                feats = self._mini_pointnet(self.mlps_before[i], side_rows[i], g2)    # (B*K, 128)
                feats = feats.view(B, K, -1).transpose(1, 2)
                feats = torch.cat((feats, dist_feature[i]), dim=1)
                return self._head(self.mlps_head[i], feats)
            bbox_feats = self._mini_pointnet(self.mlps_before[6], bbox_rows, g3)
            bbox_feats = bbox_feats.view(B, K, -1).transpose(1, 2)
            return self._head(self.mlps_head[6], bbox_feats).transpose(2, 1)

        # tensors made on the caller's stream and read inside a branch (also by the branch's backward)
        tens = lambda r: r.tensors() if isinstance(r, _GridSource) else [r]   # noqa: E731
        shared = [tens(side_rows[i]) + [dist_feature] for i in range(6)] + [tens(bbox_rows)]
        outs = self._run_branches(branch, 7, size, shared)
        end_points[f"{prefix}side_scores"] = torch.stack(outs[:6], 0)
        end_points[f"{prefix}iou_scores"] = outs[6]
        return end_points

    def _run_branches(self, branch, n, like, shared=None):
        """The seven MiniPointNet + head chains are independent: on CUDA they run as forked branches
        (branches.py), so the dozens of small kernels of one chain overlap the large GEMMs of another."""
        from .branches import run_branches
        width = int(os.environ.get("NESIE_SIDEPOOL_STREAMS", "7"))
        return run_branches([lambda i=i: branch(i) for i in range(n)], like, shared, width)
