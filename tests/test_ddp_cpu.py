"""FlatGradDDP on one process: gradients gathered into the flat buffer, the flat parameter view, and
the flat clip / AdamW / EMA path against their per-tensor torch equivalents."""
import copy

import torch
import torch.nn as nn

from nesie_b200.ddp import FlatGradDDP


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(7, 13), nn.Tanh(), nn.ReLU(), nn.Linear(13, 5), nn.ReLU(),
                         nn.Linear(5, 3))   # (no BatchNorm: a bias in front of it has a rounding-noise gradient)


def test_flat_step_equals_per_tensor_step():
    ref = _model()
    flat = copy.deepcopy(ref)
    ddp = FlatGradDDP(flat, bucket_bytes=128, flatten_parameters=True)
    assert len(ddp.buckets) >= 2
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=0.008, weight_decay=0.01)
    opt_flat = torch.optim.AdamW([ddp.flat_parameter()], lr=0.008, weight_decay=0.01)
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        x = torch.randn(32, 7, generator=g) * 5
        opt_ref.zero_grad(set_to_none=True)
        (ref(x) ** 2).sum().backward()
        total_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), 10.0)
        opt_ref.step()
        ddp.zero_grad()
        (flat(x) ** 2).sum().backward()
        ddp.finish()
        for p in ddp.params:                      # .grad is the view of the flat buffer again
            assert p.grad.data_ptr() == ddp._view[id(p)].data_ptr()
        total_flat = ddp.clip_grad_norm_(10.0)
        opt_flat.step()
        assert torch.allclose(total_ref, total_flat, rtol=1e-6)
        for a, b in zip(ref.parameters(), flat.parameters()):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), step
    # parameters without a gradient this step keep a zero gradient and are still exchanged
    ddp.zero_grad()
    (flat[0](torch.randn(4, 7)) ** 2).sum().backward()
    ddp.finish()
    assert float(flat[5].weight.grad.abs().sum()) == 0.0 and float(flat[0].weight.grad.abs().sum()) > 0.0
    # the padding between tensors stays zero
    used = torch.zeros_like(ddp.flat, dtype=torch.bool)
    for p in ddp.params:
        off = ddp.offsets()[id(p)]
        used[off:off + p.numel()] = True
    assert float(ddp.flat_params[~used].abs().sum()) == 0.0
