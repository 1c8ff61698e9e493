"""Kernel-time table of one VoteNet harness train step (torch.profiler, CUDA activities)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nesie_b200.synthetic import make_batch  # noqa: E402
from nesie_b200.votenet import VoteNetHarness  # noqa: E402

tf32 = "--tf32" in sys.argv
torch.backends.cuda.matmul.allow_tf32 = tf32
torch.backends.cudnn.allow_tf32 = tf32
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=120, max_name_column_width=70))
