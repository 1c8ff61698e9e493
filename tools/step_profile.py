"""Per-kernel GPU time of one train step of a bench.py workload (torch.profiler, CUDA kernels only).
    python tools/step_profile.py [pretrain|mean_teacher|stress|pretrain_conv] > profiles/rNN_step_profile.md"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "pretrain"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
dev = torch.device("cuda:0")
model = bench.build_model(wl).to(dev)
if wl == "mean_teacher":
    model.init_teacher()
opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01, fused=True)
inp = {k: v.to(dev) for k, v in bench.make_host_batch(wl, 0).items()}
S = bench.WORKLOADS[wl]["scenes"]
static = dict(sup_index=torch.arange(S // 2, device=dev), unsup_index=torch.arange(S // 2, S, device=dev))
it = [0]


def step():
    opt.zero_grad(set_to_none=True)
    loss = bench.step_loss(wl, model, inp, static=static)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    opt.step()
    if wl == "mean_teacher":
        model.after_train_iter(10 + it[0])
    it[0] += 1


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
print(f"workload {wl}: total kernel time per step {total / 1000:.2f} ms over "
      f"{sum(r[2] for r in rows):.0f} launches (eager, FPS chain inside the step)")
print("| kernel | us/step | launches/step | share |\n|---|---:|---:|---:|")
for k, us, n in rows[:45]:
    print(f"| `{k[:110]}` | {us:.0f} | {n:.0f} | {100 * us / total:.1f}% |")

if os.environ.get("NESIE_PROFILE_ATEN"):
    # which ATen ops (by input shape) the library kernels of the step come from
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof2:
        step()
        torch.cuda.synchronize()
    ev = [e for e in prof2.key_averages(group_by_input_shape=True)
          if e.key.startswith("aten::") and e.self_device_time_total > 0]
    ev.sort(key=lambda e: -e.self_device_time_total)
    print("\n| aten op | input shapes | us/step | calls |\n|---|---|---:|---:|")
    for e in ev[:40]:
        print(f"| `{e.key}` | `{str(e.input_shapes)[:120]}` | {e.self_device_time_total:.0f} | {e.count} |")
    tot_n = sum(e.count for e in ev)
    tot_t = sum(e.self_device_time_total for e in ev)
    print(f"\nATen ops with device time: {tot_n} calls, {tot_t / 1000:.2f} ms per step; by call count:")
    by = {}
    for e in ev:
        k = (e.key, str(e.input_shapes)[:70])
        by[k] = by.get(k, 0) + e.count
    for (k, shp), n in sorted(by.items(), key=lambda kv: -kv[1])[:45]:
        print(f"  {n:4d}  {k}  {shp}")
