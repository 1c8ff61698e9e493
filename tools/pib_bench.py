"""points_in_boxes_batch at the target-assignment shape (8 scenes x 40000 points x 64 GT boxes):
this repo's kernel vs the reference kernel compiled unmodified (oracle/_ref), CUDA events."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nesie_b200 as nb
from oracle import ref_cuda

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

g = torch.Generator().manual_seed(0)
B, M, T = 8, 40000, 64
pts = (torch.rand(B, M, 3, generator=g) * 8 - 4).cuda()
boxes = torch.cat([torch.rand(B, T, 3, generator=g) * 6 - 3, torch.rand(B, T, 3, generator=g) * 2 + 0.2,
                   torch.zeros(B, T, 1)], -1).cuda()
mine = t(lambda: nb.points_in_boxes_batch(pts, boxes))
ref = t(lambda: ref_cuda.points_in_boxes_batch(pts, boxes)) if ref_cuda.pib_available() else None
out_bytes = B * M * T * 4
print(json.dumps({"op": f"points_in_boxes_batch {B}x{M}x{T}", "this_repo_ms": round(mine, 4),
                  "reference_kernel_ms": None if ref is None else round(ref, 4),
                  "speed_up": None if ref is None else round(ref / mine, 2),
                  "out_GBps": round(out_bytes / mine / 1e6, 1)}))
