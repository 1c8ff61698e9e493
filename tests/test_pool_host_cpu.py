"""Host-side logic of the restructured MiniPointNet / SA paths (no GPU, no kernel calls): which shapes
take the pooled-epilogue / commuted-first-layer kernels and which fall back to the step-by-step ones."""
import pytest
import torch
from torch import nn

from nesie_b200 import gather_linear, mlp_rows, pool_rows


def test_pool_unit_covers_the_model_group_sizes():
    # SA levels: 64 / 32 / 16 / 16 neighbours; SidePooling: 16 grid points per face, 64 per box
    assert [pool_rows.pool_unit(k) for k in (16, 32, 64)] == [16, 32, 32]
    assert pool_rows.pool_unit(96) == 32 and pool_rows.pool_unit(224) == 32
    for k in (1, 5, 8, 24, 48, 255, 256):
        assert pool_rows.pool_unit(k) == 0          # step-by-step kernels
    assert [mlp_rows._pool_unit(k) for k in (16, 32, 64, 5)] == [16, 32, 32, 0]


def test_gather_linear_supported_widths():
    assert all(gather_linear.supported(c) for c in (4, 8, 64, 128, 256, 1024))
    assert not any(gather_linear.supported(c) for c in (0, 3, 12, 96, 260, 2048))


def _layers(chs, training=True, affine=True):
    out = []
    for a, b in zip(chs[:-1], chs[1:]):
        bn = nn.BatchNorm2d(b, affine=affine)
        bn.train(training)
        out.append((torch.zeros(b, a), bn))
    return out


def test_supported_tail_rules(monkeypatch):
    assert mlp_rows.supported_tail(_layers([131, 128, 128, 256]))
    assert not mlp_rows.supported_tail(_layers([131, 128, 128, 256], training=False))   # eval: running stats
    assert not mlp_rows.supported_tail(_layers([131, 128, 128, 256], affine=False))
    assert not mlp_rows.supported_tail(_layers([131, 128, 130, 256]))                   # width % 4
    assert not mlp_rows.supported_tail(_layers([131, 128, 128, 512]))                   # > 256 outputs
    bad = _layers([131, 128, 128, 256])
    bad[2] = (torch.zeros(256, 64), bad[2][1])                                          # channel mismatch
    assert not mlp_rows.supported_tail(bad)
    monkeypatch.setenv("NESIE_ROWS_FUSE", "0")
    assert not mlp_rows.supported_tail(_layers([131, 128, 128, 256]))


def test_switches_default_to_the_measured_choice(monkeypatch):
    for name in ("NESIE_POOL_FUSE", "NESIE_POOL_DGRAD", "NESIE_WGRAD_FORK"):
        monkeypatch.delenv(name, raising=False)
    assert pool_rows.enabled() and mlp_rows.pooled_epilogue_enabled()
    assert not pool_rows._sparse_dgrad()             # 87 us against 44 us: off
    monkeypatch.setenv("NESIE_POOL_FUSE", "0")
    assert not pool_rows.enabled() and not mlp_rows.pooled_epilogue_enabled()


def test_ops_refuse_cpu_tensors():
    t = torch.zeros(1, 8, 4)
    idx = torch.zeros(1, 2, 1, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        gather_linear.gather_linear(t, idx)
    x = torch.zeros(4, 4)
    assert pool_rows.add_bias_rows(x, torch.ones(4)).sum() == 16      # plain torch on CPU tensors (oracle twins)
