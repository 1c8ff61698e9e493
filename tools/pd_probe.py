import sys, torch
sys.path.insert(0, '/root/repo')
from nesie_b200 import _lib
torch.manual_seed(0)
G, k, N, K = 4096, 16, 128, 256
d = torch.randn(G, N, device="cuda")
arg = torch.randint(0, k, (G, N), device="cuda", dtype=torch.uint8)
w = torch.randn(N, K, device="cuda")
out = torch.empty(G * k, K, device="cuda")
y = torch.randn(G * k, K, device="cuda")
sc = torch.rand(K, device="cuda"); sh = torch.randn(K, device="cuda")
parts = torch.empty((_lib.lib().nesie_pool_wgrad_parts(G), N, K), device="cuda")
for _ in range(3):
    _lib.call("nesie_pool_dgrad", G, k, N, K, _lib.ptr(d), _lib.ptr(arg), _lib.ptr(w), _lib.ptr(out), _lib.stream())
    _lib.call("nesie_pool_wgrad", G, k, N, K, _lib.ptr(d), _lib.ptr(arg), _lib.ptr(y), _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(parts), _lib.stream())
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
_lib.call("nesie_pool_dgrad", G, k, N, K, _lib.ptr(d), _lib.ptr(arg), _lib.ptr(w), _lib.ptr(out), _lib.stream())
e1.record()
_lib.call("nesie_pool_wgrad", G, k, N, K, _lib.ptr(d), _lib.ptr(arg), _lib.ptr(y), _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(parts), _lib.stream())
e2.record()
torch.cuda.synchronize()
print("dgrad us", e0.elapsed_time(e1) * 1000, "wgrad us", e1.elapsed_time(e2) * 1000)
