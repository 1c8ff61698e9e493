"""FPS (side stream) co-running with the row GEMM (main stream): do they share SMs, and at what cost?
    NESIE_GEMM_REGS / NESIE_GEMM_SMEM_KB select the GEMM build / shared-memory budget."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nesie_b200 as nb
from nesie_b200 import _lib
from nesie_b200 import linear_rows as lr
from nesie_b200.synthetic import make_batch

pts, _, _ = make_batch(8, 40000)
xyz = pts[..., :3].contiguous().cuda()
R, K, N = 8 * 2048 * 64, 64, 128
a = torch.randn(R, K, device="cuda"); w = torch.randn(N, K, device="cuda"); out = torch.empty(R, N, device="cuda")
img = lr._pack(w, N, K, K, 1)
def gemm(n):
    for _ in range(n):
        _lib.call("nesie_gemm_nt_3xtf32", R, N, K, _lib.ptr(a), K, _lib.ptr(img), _lib.ptr(out), N, _lib.stream())
side = torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(2):
    nb.furthest_point_sample(xyz, 2048); gemm(3)
torch.cuda.synchronize()
NG = 14   # ~2 ms of GEMM work
e = [ev() for _ in range(8)]
e[0].record(); gemm(NG); e[1].record(); torch.cuda.synchronize()
e[2].record(); nb.furthest_point_sample(xyz, 2048); e[3].record(); torch.cuda.synchronize()
# concurrent
start = ev(); start.record()
side.wait_event(start)
with torch.cuda.stream(side):
    e[4].record(side); nb.furthest_point_sample(xyz, 2048); e[5].record(side)
e[6].record(); gemm(NG); e[7].record()
torch.cuda.synchronize()
print(f"regs={os.environ.get('NESIE_GEMM_REGS','96')} smem={os.environ.get('NESIE_GEMM_SMEM_KB','224')}: "
      f"gemm alone {e[0].elapsed_time(e[1]):.2f} ms | fps alone {e[2].elapsed_time(e[3]):.2f} ms | together: gemm {e[6].elapsed_time(e[7]):.2f} ms, "
      f"fps {e[4].elapsed_time(e[5]):.2f} ms, span {start.elapsed_time(e[7]):.2f}/{start.elapsed_time(e[5]):.2f} ms")
