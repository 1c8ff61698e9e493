"""PointSAModule(MSG) / PointFPModule / build_sa_module.

Host-side mirror of the reference's mmdet3d/ops/pointnet_modules/ (point_sa_module.py:11-341,
point_fp_module.py:10-78, builder.py:3-38).  Constructor arguments, forward signatures, return
values and parameter names (`mlps.{i}.layer{j}.conv.weight`, `...bn.*`) are the reference's, so a
reference state_dict loads unchanged.  mmcv is not a dependency: `ConvModule` below is the
(1x1 conv -> norm -> ReLU) block the reference builds through mmcv.cnn.ConvModule.
"""
import os
from typing import List

import torch
from torch import nn as nn
from torch.nn import functional as F

from .furthest_point_sample import Points_Sampler
from .gather_points import gather_points
from .group_points import GroupAll, QueryAndGroup
from .interpolate import three_interpolate, three_nn
from . import sa_fused
from .linear_rows import linear_rows
from . import bn_rows
from . import gather_linear
from . import mlp_rows
from .ball_query import ball_query

def _rows_linear(x, w):
    if os.environ.get('NESIE_ROWS_GEMM', 'tcgen05') == 'cublas':
        return F.linear(x, w)
    return linear_rows(x, w)


def _fused_bn():
    return os.environ.get('NESIE_ROWS_BN', 'fused') != 'aten'


_NORMS = {'BN': nn.BatchNorm2d, 'BN1d': nn.BatchNorm1d, 'BN2d': nn.BatchNorm2d}
_CONVS = {'Conv1d': nn.Conv1d, 'Conv2d': nn.Conv2d}


class ConvModule(nn.Module):
    """conv -> norm -> ReLU(inplace) with mmcv's attribute names (`conv`, `bn`, `activate`),
    `bias='auto'` rule (no conv bias when a norm layer follows) and default init
    (kaiming-normal fan_out/relu for the conv, norm weight 1 / bias 0)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1,
                 conv_cfg=dict(type='Conv2d'), norm_cfg=dict(type='BN2d'),
                 act_cfg=dict(type='ReLU'), bias='auto', padding=0, inplace=True):
        super().__init__()
        assert padding == 0, "the hot path only has 1x1 convolutions"

        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        conv_type = (conv_cfg or dict(type='Conv2d'))['type']
        self.conv = _CONVS[conv_type](in_channels, out_channels, kernel_size, stride=stride,
                                      bias=bias)
        if self.with_norm:
            self.bn = _NORMS[norm_cfg['type']](out_channels, eps=norm_cfg.get('eps', 1e-5),
                                               momentum=norm_cfg.get('momentum', 0.1))
        if self.with_activation:
            self.activate = nn.ReLU(inplace=True)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)
        if self.with_norm:
            nn.init.constant_(self.bn.weight, 1)
            nn.init.constant_(self.bn.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        if self.with_activation:
            x = self.activate(x)
        return x


class BasePointSAModule(nn.Module):
    """sample -> group -> shared MLP -> pool (point_sa_module.py:11-211)."""

    def __init__(self, num_point, radii, sample_nums, mlp_channels, fps_mod=['D-FPS'],
                 fps_sample_range_list=[-1], dilated_group=False, use_xyz=True, pool_mod='max',
                 normalize_xyz=False, grouper_return_grouped_xyz=False,
                 grouper_return_grouped_idx=False):
        super().__init__()
        assert len(radii) == len(sample_nums) == len(mlp_channels)
        assert pool_mod in ['max', 'avg']
        assert isinstance(fps_mod, (list, tuple))
        assert isinstance(fps_sample_range_list, (list, tuple))
        assert len(fps_mod) == len(fps_sample_range_list)
        if isinstance(mlp_channels, tuple):
            mlp_channels = list(map(list, mlp_channels))
        self.mlp_channels = mlp_channels
        if isinstance(num_point, int):
            self.num_point = [num_point]
        elif isinstance(num_point, (list, tuple)):
            self.num_point = num_point
        else:
            raise NotImplementedError('Error type of num_point!')
        self.pool_mod = pool_mod
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        self.fps_mod_list = fps_mod
        self.fps_sample_range_list = fps_sample_range_list
        self.points_sampler = Points_Sampler(self.num_point, self.fps_mod_list,
                                             self.fps_sample_range_list)
        for i in range(len(radii)):
            if num_point is not None:
                min_radius = radii[i - 1] if (dilated_group and i != 0) else 0
                grouper = QueryAndGroup(radii[i], sample_nums[i], min_radius=min_radius,
                                        use_xyz=use_xyz, normalize_xyz=normalize_xyz,
                                        return_grouped_xyz=grouper_return_grouped_xyz,
                                        return_grouped_idx=grouper_return_grouped_idx)
            else:
                grouper = GroupAll(use_xyz)
            self.groupers.append(grouper)

    def _sample_points(self, points_xyz, features, indices, target_xyz):
        """(new_xyz (B,M,3), indices (B,M)): given indices win, then target_xyz, else FPS."""
        xyz_flipped = points_xyz.transpose(1, 2).contiguous()
        if indices is not None:
            assert indices.shape[1] == self.num_point[0]
            new_xyz = gather_points(xyz_flipped, indices).transpose(1, 2).contiguous() \
                if self.num_point is not None else None
        elif target_xyz is not None:
            new_xyz = target_xyz.contiguous()
        else:
            indices = self.points_sampler(points_xyz, features)
            new_xyz = gather_points(xyz_flipped, indices).transpose(1, 2).contiguous() \
                if self.num_point is not None else None
        return new_xyz, indices

    def _pool_features(self, features):
        """(B, C, M, K) -> (B, C, M) by max or mean over K."""
        if self.pool_mod == 'max':
            new_features = F.max_pool2d(features, kernel_size=[1, features.size(3)])
        elif self.pool_mod == 'avg':
            new_features = F.avg_pool2d(features, kernel_size=[1, features.size(3)])
        else:
            raise NotImplementedError
        return new_features.squeeze(-1).contiguous()

    # ---- fused tcgen05 path (eval mode, folded BN, bf16 operands) ---------------------------
    fused_bf16 = False  # opt-in: outputs then match the fp32 path within 1e-2 (bf16 MLP mode)

    def _fused_ok(self, i, features):
        g = self.groupers[i]
        if self.training or not self.fused_bf16 or self.pool_mod != 'max' or features is None:
            return False
        if not isinstance(g, QueryAndGroup) or not g.use_xyz or len(self.mlps[i]) != 3:
            return False
        if (self.num_point[0] * g.sample_num) % 128 != 0:
            return False
        cs = [l.conv.out_channels for l in self.mlps[i]]
        return sa_fused.supported(g.sample_num, features.shape[1], *cs)

    def _fused_forward(self, i, points_xyz, new_xyz, features):
        g = self.groupers[i]
        key = tuple(int(t._version) for t in self.mlps[i].state_dict().values())
        cache = self.__dict__.setdefault('_fused_cache', {})
        if cache.get(i, (None,))[0] != key:
            cache[i] = (key, sa_fused.fold_mlp(self.mlps[i], features.shape[1]))
        idx = ball_query(g.min_radius, g.max_radius, g.sample_num, points_xyz.contiguous(),
                         new_xyz.contiguous())
        radius = g.max_radius if g.normalize_xyz else 0.0
        return sa_fused.sa_fused_forward(points_xyz, new_xyz, features, idx, radius, cache[i][1])

    # ---- shared MLP as row-major GEMMs ---------------------------------------------------------
    # The reference runs the shared MLP as Conv2d(1x1)+BN2d+ReLU on the channel-major
    # (B, C, M, K) tensor and max-pools with F.max_pool2d (point_sa_module.py:149-150,279-288).
    # A 1x1 convolution over (B, C, M, K) is the GEMM (B*M*K, C_in) x (C_in, C_out), BN2d's batch
    # statistics are the per-column statistics of that matrix and the pool is a max over K
    # consecutive rows, so the same parameters are applied here to one point-major copy of the
    # grouped tensor: cuBLAS GEMMs instead of cuDNN's slow fp32 1x1 wgrad/dgrad engines, and
    # BatchNorm / ReLU / max over contiguous channel rows.  Same arithmetic, same state_dict.
    rows_mlp = True

    def _rows_ok(self, i, grouped):
        g = self.groupers[i]
        if grouped is None and not (isinstance(g, QueryAndGroup) and g.use_xyz and
                                    not g.return_grouped_xyz and not g.return_grouped_idx):
            return False
        return ((grouped is None or grouped.is_cuda) and self.pool_mod == 'max' and
                all(isinstance(l.conv, nn.Conv2d) and l.conv.bias is None and l.with_norm and
                    isinstance(l.bn, nn.BatchNorm2d) for l in self.mlps[i]))

    # ---- first convolution commuted with the grouping ------------------------------------------
    # rows = [ (neighbour - centre) / radius | features[neighbour] ], so  rows @ W1^T =
    # (features @ W1[:, 3:]^T)[neighbour] + rel_xyz @ W1[:, :3]^T : the GEMM runs over the N source
    # points instead of the M * K grouped rows (16x .. 32x fewer at SA2-SA4) and the grouped tensor is
    # never built; gather_linear.cu gathers the transformed points, adds the coordinate term and takes
    # the BatchNorm statistics.  Needs coordinates without a gradient (the backbone's SA levels).
    def _gather_ok(self, i, points_xyz, new_xyz, features):
        layers = list(self.mlps[i])
        w0 = layers[0].conv.weight
        C = features.shape[1]
        return (os.environ.get("NESIE_GATHER_LINEAR", "1") != "0" and _fused_bn() and
                not (torch.is_grad_enabled() and (points_xyz.requires_grad or new_xyz.requires_grad)) and
                C >= 16 and C % 4 == 0 and w0.shape[1] == C + 3 and gather_linear.supported(w0.shape[0]) and
                points_xyz.dtype == torch.float32 and features.dtype == torch.float32 and
                mlp_rows.supported_tail([(l.conv.weight.flatten(1), l.bn) for l in layers]))

    def _gather_forward(self, i, points_xyz, new_xyz, features):
        g = self.groupers[i]
        layers = list(self.mlps[i])
        pairs = [(l.conv.weight.flatten(1), l.bn) for l in layers]
        B, C, N = features.shape
        M, K = new_xyz.shape[1], g.sample_num
        xyz, ctr = points_xyz.detach().contiguous(), new_xyz.detach().contiguous()
        idx = ball_query(g.min_radius, g.max_radius, K, xyz, ctr)
        w0 = pairs[0][0]
        table = features.transpose(1, 2).reshape(B * N, C)       # a view when features is point-major
        seeds = linear_rows(table.contiguous(), w0[:, 3:].contiguous())
        radius = g.max_radius if g.normalize_xyz else 0.0
        y, parts = gather_linear.gather_linear(seeds.view(B, N, -1), idx.view(B, M * K, 1), None, None,
                                               w0[:, :3], True, xyz, ctr, K, radius)
        out = mlp_rows.mlp_rows_tail(y, parts, pairs, K)
        return out.view(B, M, -1).transpose(1, 2)

    def _mlp_rows(self, i, x, B, M, K):
        layers = list(self.mlps[i])
        if _fused_bn():
            pairs = [(l.conv.weight.flatten(1), l.bn) for l in layers]
            if x.shape[1] != pairs[0][0].shape[1]:  # rows zero-padded to a multiple of 4 columns
                pairs[0] = (F.pad(pairs[0][0], (0, x.shape[1] - pairs[0][0].shape[1])), pairs[0][1])
            if mlp_rows.supported(x, pairs):
                # BatchNorm fused into the GEMMs: statistics in the epilogue, apply in the prologue
                x = mlp_rows.mlp_rows(x, pairs, K)
                return x.view(B, M, -1).transpose(1, 2)
        for li, layer in enumerate(layers):
            bn = layer.bn
            # fp32-parity GEMM on tcgen05 (3xTF32); NESIE_ROWS_GEMM=cublas selects the library GEMM
            w = layer.conv.weight.flatten(1)
            if x.shape[1] != w.shape[1]:  # grouped rows were zero-padded to a multiple of 4 columns
                w = F.pad(w, (0, x.shape[1] - w.shape[1]))
            x = _rows_linear(x, w)
            last = li == len(layers) - 1
            if _fused_bn() and bn_rows.supported(x, bn, K if last else 0):
                # fused BatchNorm(batch stats) + ReLU (+ the max-pool over the K rows of a group)
                x = bn_rows.bn_relu_rows(x, bn, K if last else 0)
                if last:
                    # (B, C, M) as the reference returns it, as a VIEW of the point-major rows: the
                    # next SA level's row gather wants exactly that storage order
                    return x.view(B, M, -1).transpose(1, 2)
                continue
            if bn.training and bn.track_running_stats:
                bn_rows.count_batch(bn)
            x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                             bn.training or not bn.track_running_stats, bn.momentum, bn.eps)
            x = F.relu(x, inplace=True)
        x = x.view(B, M, K, -1).amax(dim=2)
        return x.transpose(1, 2).contiguous()

    def forward(self, points_xyz, features=None, indices=None, target_xyz=None):
        """-> (new_xyz (B,M,3), new_features (B, sum C_out, M), indices (B,M) int32)."""
        new_features_list = []
        new_xyz, indices = self._sample_points(points_xyz, features, indices, target_xyz)
        for i in range(len(self.groupers)):
            if self._fused_ok(i, features):
                new_features_list.append(self._fused_forward(i, points_xyz, new_xyz, features))
                continue
            if self.rows_mlp and points_xyz.is_cuda and features is not None and \
                    self._rows_ok(i, None):
                g = self.groupers[i]
                if self._gather_ok(i, points_xyz, new_xyz, features):
                    new_features_list.append(self._gather_forward(i, points_xyz, new_xyz, features))
                    continue
                rows = g.forward_rows(points_xyz, new_xyz, features, pad_to=4)
                new_features_list.append(self._mlp_rows(i, rows, points_xyz.shape[0],
                                                        new_xyz.shape[1], g.sample_num))
                continue
            grouped_results = self.groupers[i](points_xyz, new_xyz, features)
            if self.rows_mlp and self._rows_ok(i, grouped_results):
                B_, C0_, M_, K_ = grouped_results.shape
                rows = grouped_results.permute(0, 2, 3, 1).reshape(B_ * M_ * K_, C0_)
                new_features = self._mlp_rows(i, rows, B_, M_, K_)
            else:
                new_features = self.mlps[i](grouped_results)
                new_features = self._pool_features(new_features)
            new_features_list.append(new_features)
        new_features = new_features_list[0] if len(new_features_list) == 1 \
            else torch.cat(new_features_list, dim=1)
        return new_xyz, new_features, indices


class PointSAModuleMSG(BasePointSAModule):
    """Multi-scale grouping SA module (point_sa_module.py:214-289)."""

    def __init__(self, num_point, radii, sample_nums, mlp_channels, fps_mod=['D-FPS'],
                 fps_sample_range_list=[-1], dilated_group=False, norm_cfg=dict(type='BN2d'),
                 use_xyz=True, pool_mod='max', normalize_xyz=False, bias='auto'):
        super().__init__(num_point=num_point, radii=radii, sample_nums=sample_nums,
                         mlp_channels=mlp_channels, fps_mod=fps_mod,
                         fps_sample_range_list=fps_sample_range_list,
                         dilated_group=dilated_group, use_xyz=use_xyz, pool_mod=pool_mod,
                         normalize_xyz=normalize_xyz)
        for i in range(len(self.mlp_channels)):
            mlp_channel = self.mlp_channels[i]
            if use_xyz:
                mlp_channel[0] += 3
            mlp = nn.Sequential()
            for j in range(len(mlp_channel) - 1):
                mlp.add_module(f'layer{j}',
                               ConvModule(mlp_channel[j], mlp_channel[j + 1], kernel_size=(1, 1),
                                          stride=(1, 1), conv_cfg=dict(type='Conv2d'),
                                          norm_cfg=norm_cfg, bias=bias))
            self.mlps.append(mlp)


class PointSAModule(PointSAModuleMSG):
    """Single-scale grouping SA module (point_sa_module.py:292-341)."""

    def __init__(self, mlp_channels, num_point=None, radius=None, num_sample=None,
                 norm_cfg=dict(type='BN2d'), use_xyz=True, pool_mod='max', fps_mod=['D-FPS'],
                 fps_sample_range_list=[-1], normalize_xyz=False):
        super().__init__(mlp_channels=[list(mlp_channels)], num_point=num_point, radii=[radius],
                         sample_nums=[num_sample], norm_cfg=norm_cfg, use_xyz=use_xyz,
                         pool_mod=pool_mod, fps_mod=fps_mod,
                         fps_sample_range_list=fps_sample_range_list,
                         normalize_xyz=normalize_xyz)


class PointFPModule(nn.Module):
    """Feature propagation: three_nn -> inverse-distance weights -> three_interpolate ->
    concat skip -> shared MLP (point_fp_module.py:10-78)."""

    def __init__(self, mlp_channels: List[int], norm_cfg: dict = dict(type='BN2d'), init_cfg=None):
        super().__init__()
        self.fp16_enabled = False
        self.mlps = nn.Sequential()
        for i in range(len(mlp_channels) - 1):
            self.mlps.add_module(f'layer{i}',
                                 ConvModule(mlp_channels[i], mlp_channels[i + 1],
                                            kernel_size=(1, 1), stride=(1, 1),
                                            conv_cfg=dict(type='Conv2d'), norm_cfg=norm_cfg))

    def forward(self, target, source, target_feats, source_feats):
        """target (B,n,3), source (B,m,3), target_feats (B,C1,n), source_feats (B,C2,m)
        -> (B, mlp[-1], n)."""
        if source is not None:
            dist, idx = three_nn(target.contiguous(), source.contiguous())
            dist_reciprocal = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_reciprocal, dim=2, keepdim=True)
            weight = dist_reciprocal / norm
            interpolated_feats = three_interpolate(source_feats.contiguous(), idx, weight)
        else:
            interpolated_feats = source_feats.expand(*source_feats.size()[0:2], target.size(1))
        if target_feats is not None:
            new_features = torch.cat([interpolated_feats, target_feats], dim=1)
        else:
            new_features = interpolated_feats
        if self.rows_mlp and new_features.is_cuda:
            return self._mlp_rows(new_features)
        new_features = self.mlps(new_features.unsqueeze(-1))
        return new_features.squeeze(-1)

    rows_mlp = True  # see BasePointSAModule._mlp_rows: same parameters, GEMM formulation

    def _mlp_rows(self, feats):
        B, C, n = feats.shape
        x = feats.transpose(1, 2).reshape(B * n, C)
        if _fused_bn() and all(l.conv.bias is None for l in self.mlps):
            pairs = [(l.conv.weight.flatten(1), l.bn) for l in self.mlps]
            if mlp_rows.supported(x, pairs):
                return mlp_rows.mlp_rows(x, pairs).view(B, n, -1).transpose(1, 2)
        for layer in self.mlps:
            bn = layer.bn
            x = _rows_linear(x, layer.conv.weight.flatten(1))
            if layer.conv.bias is not None:
                x = x + layer.conv.bias
            if _fused_bn() and bn_rows.supported(x, bn):
                x = bn_rows.bn_relu_rows(x, bn)
                continue
            if bn.training and bn.track_running_stats:
                bn_rows.count_batch(bn)
            x = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                             bn.training or not bn.track_running_stats, bn.momentum, bn.eps)
            x = F.relu(x, inplace=True)
        return x.view(B, n, -1).transpose(1, 2)  # (B, C, n) view of point-major rows


SA_MODULES = {'PointSAModule': PointSAModule, 'PointSAModuleMSG': PointSAModuleMSG}


def build_sa_module(cfg, *args, **kwargs):
    """Instantiate an SA module from a {type: ..., **module_args} dict (builder.py:6-38)."""
    if cfg is None:
        cfg_ = dict(type='PointSAModule')
    else:
        if not isinstance(cfg, dict):
            raise TypeError('cfg must be a dict')
        if 'type' not in cfg:
            raise KeyError('the cfg dict must contain the key "type"')
        cfg_ = cfg.copy()
    module_type = cfg_.pop('type')
    if module_type not in SA_MODULES:
        raise KeyError(f'Unrecognized module type {module_type}')
    return SA_MODULES[module_type](*args, **kwargs, **cfg_)
