"""PointNet2SASSG backbone (4 x SA + 2 x FP for VoteNet).

Mirror of mmdet3d/models/backbones/pointnet2_sa_ssg.py:11-142 and base_pointnet.py:19-37: same
constructor, same output dict (fp_xyz / fp_features / fp_indices / sa_xyz / sa_features /
sa_indices) and the same parameter names (SA_modules.{i}.mlps.0.layer{j}..., FP_modules.{i}...).

B200-first scheduling: furthest point sampling depends only on coordinates, never on features,
so the whole FPS chain (40000 -> 2048 -> 1024 -> 512 -> 256) is issued up front on a side
stream and overlaps the ball queries / MLPs of the earlier levels on the main stream; each SA
module then receives its `indices` (the reference API already accepts them,
point_sa_module.py:160-165).  FPS is a serial, latency-bound kernel that occupies at most
b x 16 SMs, so the remaining SMs do the dense work meanwhile.
"""
import torch
from torch import nn as nn

from .furthest_point_sample import furthest_point_sample
from .gather_points import gather_points
from .pointnet_modules import PointFPModule, build_sa_module


class PointNet2SASSG(nn.Module):

    def __init__(self, in_channels, num_points=(2048, 1024, 512, 256),
                 radius=(0.2, 0.4, 0.8, 1.2), num_samples=(64, 32, 16, 16),
                 sa_channels=((64, 64, 128), (128, 128, 256), (128, 128, 256), (128, 128, 256)),
                 fp_channels=((256, 256), (256, 256)), norm_cfg=dict(type='BN2d'),
                 sa_cfg=dict(type='PointSAModule', pool_mod='max', use_xyz=True,
                             normalize_xyz=True),
                 init_cfg=None, overlap_fps=True):
        super().__init__()
        self.num_sa = len(sa_channels)
        self.num_fp = len(fp_channels)
        self.num_points = tuple(num_points)
        self.overlap_fps = overlap_fps
        assert len(num_points) == len(radius) == len(num_samples) == len(sa_channels)
        assert len(sa_channels) >= len(fp_channels)

        self.SA_modules = nn.ModuleList()
        sa_in_channel = in_channels - 3
        skip_channel_list = [sa_in_channel]
        for sa_index in range(self.num_sa):
            cur_sa_mlps = [sa_in_channel] + list(sa_channels[sa_index])
            sa_out_channel = cur_sa_mlps[-1]
            self.SA_modules.append(
                build_sa_module(num_point=num_points[sa_index], radius=radius[sa_index],
                                num_sample=num_samples[sa_index], mlp_channels=cur_sa_mlps,
                                norm_cfg=norm_cfg, cfg=sa_cfg))
            skip_channel_list.append(sa_out_channel)
            sa_in_channel = sa_out_channel

        self.FP_modules = nn.ModuleList()
        fp_source_channel = skip_channel_list.pop()
        fp_target_channel = skip_channel_list.pop()
        for fp_index in range(len(fp_channels)):
            cur_fp_mlps = [fp_source_channel + fp_target_channel] + list(fp_channels[fp_index])
            self.FP_modules.append(PointFPModule(mlp_channels=cur_fp_mlps))
            if fp_index != len(fp_channels) - 1:
                fp_source_channel = cur_fp_mlps[-1]
                fp_target_channel = skip_channel_list.pop()
        self._fps_stream = None

    @staticmethod
    def _split_point_feats(points):
        """(B,N,3+C) -> xyz (B,N,3), features (B,C,N) or None (base_pointnet.py:19-37)."""
        xyz = points[..., 0:3].contiguous()
        if points.size(-1) > 3:
            features = points[..., 3:].transpose(1, 2).contiguous()
        else:
            features = None
        return xyz, features

    def _fps_chain(self, xyz):
        """All SA levels' FPS indices, issued on a side stream.  Returns [(indices, event)]."""
        if self._fps_stream is None:
            self._fps_stream = torch.cuda.Stream(device=xyz.device)
        main = torch.cuda.current_stream(xyz.device)
        side = self._fps_stream
        side.wait_stream(main)
        out = []
        with torch.cuda.stream(side), torch.no_grad():
            cur = xyz
            for i in range(self.num_sa):
                idx = furthest_point_sample(cur, self.num_points[i])
                ev = torch.cuda.Event()
                ev.record(side)
                out.append((idx, ev))
                if i + 1 < self.num_sa:
                    cur = gather_points(cur.transpose(1, 2).contiguous(), idx) \
                        .transpose(1, 2).contiguous()
                idx.record_stream(main)
        return out

    def fps_chain(self, points, given=(), stop=None):
        """The SA levels' FPS indices [(B, num_points[i]) int32] for `points`, on the current
        stream.  They depend on coordinates only, so an input pipeline can compute them for batch
        t+1 while batch t trains and hand them to forward(points, fps_indices=...).
        given: indices of the first len(given) levels computed earlier (they are only gathered);
        stop: number of levels to produce in total (default all).  Returns levels len(given)..stop-1,
        which lets a pipeline run the long first level and the short remaining ones in separate
        time slots."""
        stop = self.num_sa if stop is None else stop
        cur = points[..., 0:3].contiguous()
        out = []
        with torch.no_grad():
            for i in range(stop):
                if i < len(given):
                    idx = given[i]
                else:
                    idx = furthest_point_sample(cur, self.num_points[i])
                    out.append(idx)
                if i + 1 < stop:
                    cur = gather_points(cur.transpose(1, 2).contiguous(), idx) \
                        .transpose(1, 2).contiguous()
        return out

    def forward(self, points, fps_indices=None, after_level=None):
        """points (B, N, 3 + input_feature_dim) -> dict of lists, as the reference returns.
        fps_indices: optional precomputed result of fps_chain(points).
        after_level: optional callable(i) invoked once SA level i has been issued (scheduling hook:
        bench.py forks the next batch's FPS chain there, away from the large SA1 / SA2 GEMMs)."""
        xyz, features = self._split_point_feats(points)
        batch, num_points = xyz.shape[:2]
        indices = torch.arange(num_points, device=xyz.device, dtype=torch.long) \
            .unsqueeze(0).repeat(batch, 1)

        sa_xyz, sa_features, sa_indices = [xyz], [features], [indices]
        use_chain = self.overlap_fps and all(
            getattr(m, 'fps_mod_list', None) == ['D-FPS'] and
            list(getattr(m, 'fps_sample_range_list', [])) == [-1] for m in self.SA_modules)
        chain = self._fps_chain(xyz) if (use_chain and fps_indices is None) else None
        for i in range(self.num_sa):
            pre = fps_indices[i] if fps_indices is not None else None
            if chain is not None:
                pre, ev = chain[i]
                torch.cuda.current_stream(xyz.device).wait_event(ev)
            cur_xyz, cur_features, cur_indices = self.SA_modules[i](sa_xyz[i], sa_features[i],
                                                                     indices=pre)
            sa_xyz.append(cur_xyz)
            sa_features.append(cur_features)
            sa_indices.append(torch.gather(sa_indices[-1], 1, cur_indices.long()))
            if after_level is not None:
                after_level(i)

        fp_xyz, fp_features, fp_indices = [sa_xyz[-1]], [sa_features[-1]], [sa_indices[-1]]
        for i in range(self.num_fp):
            fp_features.append(self.FP_modules[i](sa_xyz[self.num_sa - i - 1],
                                                  sa_xyz[self.num_sa - i],
                                                  sa_features[self.num_sa - i - 1],
                                                  fp_features[-1]))
            fp_xyz.append(sa_xyz[self.num_sa - i - 1])
            fp_indices.append(sa_indices[self.num_sa - i - 1])
        return dict(fp_xyz=fp_xyz, fp_features=fp_features, fp_indices=fp_indices,
                    sa_xyz=sa_xyz, sa_features=sa_features, sa_indices=sa_indices)
