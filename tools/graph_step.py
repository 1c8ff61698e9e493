"""Experiment: the VoteNet train step eager vs captured in one CUDA graph (static input buffers)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesie_b200.synthetic import make_batch  # noqa: E402
from nesie_b200.votenet import VoteNetHarness  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
model = VoteNetHarness().cuda()
opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01, fused=True, capturable=True)
pts, gb, gl = make_batch(8, 40000)
pts = pts.cuda()
gb = [b.cuda() for b in gb]
gl = [l.cuda() for l in gl]
boxes, labels, valid = model._pad_gt(gb, gl, pts.device)


def step():
    opt.zero_grad(set_to_none=False)
    loss, _ = model.train_step_loss(pts, gb, gl)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    opt.step()
    return loss


def timeit(fn, n=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for _ in range(3):
    step()
print("eager ms/step", round(timeit(step), 3))
t0 = time.perf_counter()
for _ in range(5):
    step()
t_cpu = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
print("eager CPU issue ms/step (no sync)", round(t_cpu, 3))
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        static_loss = step()
    print("graph ms/step", round(timeit(g.replay), 3), "loss", float(static_loss))
except Exception as e:  # noqa: BLE001
    print("graph capture failed:", repr(e)[:400])
