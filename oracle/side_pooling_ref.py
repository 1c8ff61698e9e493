"""CPU restatement of the reference SidePooling forward (TEST INFRASTRUCTURE ONLY).

`SidePoolingOracle` shares parameters, geometry code and call structure with
nesie_b200.side_pooling.SidePooling and replaces its three hot-path hooks with plain CPU torch that
follows the reference's own formulation (mmdet3d/models/dense_heads/side_pooling_module.py):
  _grid_rows      :183-243  three nearest seeds (oracle three_nn, the C restatement of the kernel
                            mmcv.ops.three_nn wraps), distances RE-computed from the gathered points
                            as sqrt(sum(d*d)), weights 1/(d+1e-8) normalised, weighted sum of the
                            three gathered feature rows, relative grid coordinates in front
  _mini_pointnet  :360-370  Conv2d / BatchNorm2d / ReLU modules on a (1, C, boxes, G) tensor
  _head           :54-80    the nn.Sequential itself
Pinned against outputs of the reference class itself (tests/golden/make_golden_sidepool.py)."""
import torch

from nesie_b200.side_pooling import SidePooling
from oracle import cpu


class SidePoolingOracle(SidePooling):

    def _grid_rows(self, origin_xyz, origin_features, grid, center, nn_cache=None):
        B, T = grid.shape[:2]
        K = center.shape[1]
        C = origin_features.shape[1]
        _, idx = cpu.three_nn(grid, origin_xyz)                       # (B, T, 3) int32
        idx = idx.long()
        near = torch.gather(origin_xyz, 1, idx.view(B, -1, 1).expand(-1, -1, 3))   # (B, T*3, 3)
        d = near - grid.unsqueeze(2).expand(-1, -1, 3, -1).reshape(B, -1, 3)
        dist = torch.sqrt(torch.sum(d * d, dim=2))
        weight = (1 / (dist + 1e-8)).view(B, -1, 3)
        weight = weight / torch.sum(weight, dim=2, keepdim=True)
        table = origin_features.transpose(1, 2)                        # (B, N, C)
        feats = torch.stack([table[b].index_select(0, idx[b].reshape(-1)) for b in range(B)], 0)
        feats = torch.sum(feats.view(B, -1, 3, C) * weight.unsqueeze(-1), dim=2)   # (B, T, C)
        head = grid.view(B, K, T // K, 3) - center.unsqueeze(2)
        return torch.cat([head.reshape(B * T, 3), feats.reshape(B * T, C)], dim=1)

    def _side_rows(self, origin_xyz, origin_features, side_grid, center, nn_cache=None):
        rows = self._grid_rows(origin_xyz, origin_features, side_grid, center)
        B, K = center.shape[:2]
        rows = rows.view(B * K, 6, -1, rows.shape[-1])
        return [rows[:, i].reshape(-1, rows.shape[-1]) for i in range(6)]

    def _mini_pointnet(self, mpn, rows, G):
        R, C = rows.shape
        points = rows.view(R // G, G, C).permute(2, 0, 1).unsqueeze(0)    # (1, C, boxes, G)
        return mpn(points)[0].transpose(0, 1)                             # (boxes, feature_dim)

    def _head(self, seq, x):
        return seq(x)
