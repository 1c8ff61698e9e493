"""Per-kernel GPU time of the SidePooling head at the Nesie shape (eval forward, train fwd+bwd)."""
import os, sys, torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nesie_b200.side_pooling import SidePooling
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
B, K, N, C, NC = 8, 512, 1024, 256, 18
m = SidePooling(NC, 1, NC, None, K // 2, "vote", seed_feat_dim=C).cuda()
g = torch.Generator().manual_seed(3)
boxes = [(torch.rand(B, K, 3, generator=g) * 6 - 3).cuda(), (torch.rand(B, K, 3, generator=g) * 1.5 + 0.2).cuda(), torch.zeros(B, K).cuda()]
ep = {"seed_points": (torch.rand(B, N, 3, generator=g) * 8 - 4).cuda(), "seed_features": torch.randn(B, C, N, generator=g).cuda(),
      "bbox_probs": torch.softmax(torch.randn(B, 6, 33, K // 2, generator=g), 2).cuda()}
for train in (False, True):
    m.train(train)
    def step():
        with (torch.enable_grad() if train else torch.no_grad()):
            out = m(*boxes, dict(ep))
            if train:
                (out["side_scores"].sum() + out["iou_scores"].sum()).backward()
    for _ in range(2): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    rows = sorted(((e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0), reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"==== {'train fwd+bwd' if train else 'eval fwd'}: {tot / 1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
    for us, n, k in rows[:14]:
        print(f"{us:8.0f} us x{n:4d}  {k[:100]}")
