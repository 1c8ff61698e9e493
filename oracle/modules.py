"""Reference-shaped forward (and backward) of the PointNet++ modules on the CPU.

TEST INFRASTRUCTURE ONLY.  Orchestration follows the reference's python
(ops/group_points/group_points.py:64-128, ops/pointnet_modules/point_sa_module.py:103-211,
point_fp_module.py:39-78, models/backbones/pointnet2_sa_ssg.py:88-142) with the oracle's C ops
in place of the CUDA extensions and torch-CPU Conv/BN for the shared MLPs.  It takes the
*parameters* from a nesie_b200 module (same nn.Module layout as the reference) so both sides
of a parity test share weights.  This is also the "reference CPU forward" bench.py times.
"""
import torch
from torch.autograd import Function
from torch.nn import functional as F

from . import cpu


class _Gather(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return cpu.gather_points(features, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return cpu.gather_points_grad(g, idx, ctx.n), None


class _Group(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return cpu.grouping_operation(features, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return cpu.grouping_operation_grad(g, idx, ctx.n), None


class _Interp(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.m = features.shape[2]
        return cpu.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, g):
        idx, weight = ctx.saved_tensors
        return cpu.three_interpolate_grad(g, idx, weight, ctx.m), None, None


gather_points = _Gather.apply
grouping_operation = _Group.apply
three_interpolate = _Interp.apply


def query_and_group(points_xyz, center_xyz, features, radius, nsample, min_radius=0.0,
                    use_xyz=True, normalize_xyz=False):
    """group_points.py:81-116.  `/= radius` is the CUDA semantics the reference runs with:
    multiplication by the fp32 reciprocal (ATen BinaryDivTrueKernel.cu, CPU-scalar divisor)."""
    idx = cpu.ball_query(min_radius, radius, nsample, points_xyz, center_xyz)
    xyz_trans = points_xyz.transpose(1, 2).contiguous()
    grouped_xyz = grouping_operation(xyz_trans, idx)
    grouped_xyz = grouped_xyz - center_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        inv = (torch.tensor(1.0, dtype=torch.float32) /
               torch.tensor(radius, dtype=torch.float32)).item()
        grouped_xyz = grouped_xyz * inv
    if features is not None:
        grouped_features = grouping_operation(features.contiguous(), idx)
        new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if use_xyz \
            else grouped_features
    else:
        new_features = grouped_xyz
    return new_features, idx


def sa_forward(module, points_xyz, features=None, indices=None, target_xyz=None):
    """BasePointSAModule.forward with `module` = a (CPU copy of a) nesie_b200 PointSAModule."""
    xyz_flipped = points_xyz.transpose(1, 2).contiguous()
    if indices is None and target_xyz is None:
        indices = cpu.furthest_point_sample(points_xyz, module.num_point[0])
    if indices is not None:
        new_xyz = gather_points(xyz_flipped, indices).transpose(1, 2).contiguous()
    else:
        new_xyz = target_xyz.contiguous()
    outs = []
    for grouper, mlp in zip(module.groupers, module.mlps):
        grouped, _ = query_and_group(points_xyz, new_xyz, features, grouper.max_radius,
                                     grouper.sample_num, grouper.min_radius, grouper.use_xyz,
                                     grouper.normalize_xyz)
        x = mlp(grouped)
        if module.pool_mod == 'max':
            x = F.max_pool2d(x, kernel_size=[1, x.size(3)])
        else:
            x = F.avg_pool2d(x, kernel_size=[1, x.size(3)])
        outs.append(x.squeeze(-1).contiguous())
    return new_xyz, torch.cat(outs, dim=1), indices


def fp_forward(module, target, source, target_feats, source_feats):
    """PointFPModule.forward (point_fp_module.py:39-78)."""
    dist, idx = cpu.three_nn(target, source)
    dist_reciprocal = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_reciprocal, dim=2, keepdim=True)
    weight = dist_reciprocal / norm
    interpolated = three_interpolate(source_feats.contiguous(), idx, weight.contiguous())
    if target_feats is not None:
        x = torch.cat([interpolated, target_feats], dim=1)
    else:
        x = interpolated
    return module.mlps(x.unsqueeze(-1)).squeeze(-1)


def backbone_forward(backbone, points):
    """PointNet2SASSG.forward (pointnet2_sa_ssg.py:88-142) with `backbone` on the CPU."""
    xyz = points[..., 0:3].contiguous()
    features = points[..., 3:].transpose(1, 2).contiguous() if points.size(-1) > 3 else None
    B, N = xyz.shape[:2]
    indices = torch.arange(N, dtype=torch.long, device=xyz.device).unsqueeze(0).repeat(B, 1)
    sa_xyz, sa_features, sa_indices = [xyz], [features], [indices]
    for i in range(backbone.num_sa):
        cx, cf, ci = sa_forward(backbone.SA_modules[i], sa_xyz[i], sa_features[i])
        sa_xyz.append(cx)
        sa_features.append(cf)
        sa_indices.append(torch.gather(sa_indices[-1], 1, ci.long()))
    fp_xyz, fp_features, fp_indices = [sa_xyz[-1]], [sa_features[-1]], [sa_indices[-1]]
    ns = backbone.num_sa
    for i in range(backbone.num_fp):
        fp_features.append(fp_forward(backbone.FP_modules[i], sa_xyz[ns - i - 1], sa_xyz[ns - i],
                                      sa_features[ns - i - 1], fp_features[-1]))
        fp_xyz.append(sa_xyz[ns - i - 1])
        fp_indices.append(sa_indices[ns - i - 1])
    return dict(fp_xyz=fp_xyz, fp_features=fp_features, fp_indices=fp_indices, sa_xyz=sa_xyz,
                sa_features=sa_features, sa_indices=sa_indices)
