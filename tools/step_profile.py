"""Per-kernel GPU time of one VoteNet harness train step (torch.profiler, CUDA kernels only)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesie_b200.synthetic import make_batch  # noqa: E402
from nesie_b200.votenet import VoteNetHarness  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
model = VoteNetHarness().cuda()
opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01, fused=True)
pts, gb, gl = make_batch(8, 40000)
pts = pts.cuda()
gb = [b.cuda() for b in gb]
gl = [l.cuda() for l in gl]


def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = model.train_step_loss(pts, gb, gl)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
NS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NS):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total / NS, e.count / NS) for e in prof.key_averages()
        if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
total = sum(r[1] for r in rows)
print(f"total kernel time per step: {total / 1000:.2f} ms over {sum(r[2] for r in rows):.0f} launches")
print("| kernel | us/step | launches/step | share |\n|---|---:|---:|---:|")
for k, us, n in rows[:40]:
    print(f"| `{k[:110]}` | {us:.0f} | {n:.0f} | {100 * us / total:.1f}% |")
