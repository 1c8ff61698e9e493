"""Host-side logic that needs no GPU: constructor contracts of the reference API mirror,
parameter naming, EMA momentum schedule, refusal of CPU tensors."""
import pytest
import torch

import nesie_b200 as nb


def test_query_and_group_constructor_contracts():
    with pytest.raises(AssertionError):
        nb.QueryAndGroup(0.2, 16, return_unique_cnt=True)  # needs uniform_sample
    with pytest.raises(NotImplementedError):
        nb.QueryAndGroup(None, 16)  # kNN grouping is off the Nesie path
    g = nb.QueryAndGroup(0.2, 16, min_radius=0.1, normalize_xyz=True)
    assert (g.max_radius, g.min_radius, g.sample_num, g.use_xyz) == (0.2, 0.1, 16, True)


def test_points_sampler_contracts():
    with pytest.raises(AssertionError):
        nb.Points_Sampler([16, 16], ['D-FPS'], [-1])
    with pytest.raises(ValueError):
        nb.Points_Sampler([16], ['X-FPS'], [-1])


def test_build_sa_module_and_names():
    with pytest.raises(KeyError):
        nb.build_sa_module(dict(type='Nope'), mlp_channels=[1, 2])
    with pytest.raises(TypeError):
        nb.build_sa_module('PointSAModule')
    m = nb.build_sa_module(dict(type='PointSAModule', pool_mod='max', use_xyz=True, normalize_xyz=True),
                           num_point=16, radius=0.2, num_sample=8, mlp_channels=[1, 8, 16],
                           norm_cfg=dict(type='BN2d'))
    assert m.mlps[0].layer0.conv.weight.shape == (8, 4, 1, 1)
    assert m.mlps[0].layer0.conv.bias is None  # bias='auto' with a norm layer
    assert isinstance(m.mlps[0].layer1.bn, torch.nn.BatchNorm2d)


def test_backbone_layout_matches_config():
    bb = nb.PointNet2SASSG(in_channels=4)
    assert [m.num_point for m in bb.SA_modules] == [[2048], [1024], [512], [256]]
    assert [m.groupers[0].sample_num for m in bb.SA_modules] == [64, 32, 16, 16]
    assert bb.SA_modules[2].mlps[0].layer0.conv.in_channels == 259
    assert bb.FP_modules[1].mlps.layer0.conv.in_channels == 512
    n = sum(p.numel() for p in bb.parameters())
    assert 0.6e6 < n < 1.2e6


def test_ops_raise_on_cpu_tensors():
    with pytest.raises(RuntimeError, match="CUDA"):
        nb.furthest_point_sample(torch.rand(1, 10, 3), 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        nb.three_nn(torch.rand(1, 4, 3), torch.rand(1, 4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        nb.aligned_3d_nms(torch.rand(3, 6), torch.rand(3), torch.zeros(3), 0.25)
    with pytest.raises(AssertionError):
        nb.ball_query(0.5, 0.2, 4, torch.rand(1, 4, 3), torch.rand(1, 2, 3))  # min >= max


def test_calc_square_dist():
    a, b = torch.rand(2, 5, 4), torch.rand(2, 7, 4)
    d = nb.calc_square_dist(a, b, norm=False)
    want = ((a[:, :, None] - b[:, None]) ** 2).sum(-1)
    assert torch.allclose(d, want, atol=1e-5)
