"""Which layout copies / fills / small reductions does one train step issue (shapes, device time)?"""
import os, sys, torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200.synthetic import make_batch
from nesie_b200.votenet import VoteNetHarness
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
model = VoteNetHarness().cuda()
opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01, fused=True)
pts, gb, gl = make_batch(8, 40000)
pts = pts.cuda(); gb = [b.cuda() for b in gb]; gl = [l.cuda() for l in gl]
def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = model.train_step_loss(pts, gb, gl)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::fill_", "aten::zero_", "aten::sum", "aten::add", "aten::add_", "aten::mul", "aten::cat", "aten::zeros"):
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
for us, n, k, shp in rows[:45]:
    print(f"{us:8.0f} us  x{n:3d}  {k:18s} {shp}")
