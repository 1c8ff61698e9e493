"""3xTF32 tcgen05 row GEMM: fp32 parity (the SA shared-MLP GEMMs in training)."""
import pytest
import torch

from nesie_b200.linear_rows import gemm_nt, linear_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,N,K", [(128, 64, 4), (1000, 64, 64), (4096, 128, 131), (300, 256, 128),
                                   (65536, 128, 259), (257, 16, 32), (5000, 4, 64), (2048, 259 - 3, 128),
                                   (1, 128, 128), (70000, 256, 128), (4096, 515, 256), (8192, 512, 256)])
def test_gemm_matches_float64(R, N, K):
    torch.manual_seed(R + N + K)
    a = torch.randn(R, K, device="cuda")
    w = torch.randn(N, K, device="cuda")
    want = (a.double() @ w.double().t())
    got = gemm_nt(a, w)
    err = ((got.double() - want).abs().max() / want.abs().max()).item()
    assert err < 6e-6, err  # fp32 parity bar is 1e-5; cuBLAS sgemm sits near 1e-6 here
    wt = w.t().contiguous()          # (K, N) storage, same product
    got_t = gemm_nt(a, wt, transpose_w=True)
    assert ((got_t.double() - want).abs().max() / want.abs().max()).item() < 6e-6


def test_linear_rows_autograd():
    torch.manual_seed(0)
    x = torch.randn(3000, 131, device="cuda", requires_grad=True)
    w = torch.randn(128, 131, device="cuda", requires_grad=True)
    y = linear_rows(x, w)
    g = torch.randn_like(y)
    y.backward(g)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yd = xd @ wd.t()
    yd.backward(g.double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()  # noqa: E731
    assert rel(y.detach(), yd.detach()) < 6e-6
    assert rel(x.grad, xd.grad) < 6e-6
    assert rel(w.grad, wd.grad) < 2e-5


@pytest.mark.parametrize("R,N,K", [(64, 128, 32), (1000, 64, 64), (4096, 128, 131), (300, 256, 128),
                                   (65536, 128, 259), (5000, 64, 4), (33, 16, 8), (70000, 256, 128),
                                   (200000, 64, 64)])
def test_wgrad_matches_float64(R, N, K):
    from nesie_b200.linear_rows import wgrad
    torch.manual_seed(R + N + K)
    gy = torch.randn(R, N, device="cuda")
    x = torch.randn(R, K, device="cuda")
    want = gy.double().t() @ x.double()
    got = wgrad(gy, x)
    err = ((got.double() - want).abs().max() / want.abs().max()).item()
    assert err < 2e-5, err  # weight gradient: in-TMEM accumulation truncates (see gemm_3xtf32.cu)


@pytest.mark.parametrize("ns,N,K", [(148, 64, 64), (1, 128, 132), (37, 256, 260), (5, 4, 4)])
def test_sum_partials_is_deterministic_and_exact_enough(ns, N, K):
    from nesie_b200.linear_rows import sum_partials
    torch.manual_seed(ns)
    parts = torch.randn(ns, N, K, device="cuda")
    got = sum_partials(parts)
    assert torch.equal(got, sum_partials(parts))                       # fixed association
    want = parts.double().sum(0)
    assert ((got.double() - want).abs().max() / want.abs().max()).item() < 2e-6


def test_cta_pair_multicast_mode_matches_single_ctas():
    """NESIE_GEMM_PAIR=1 (clusters of two CTAs sharing the weight slabs by multicast) gives the same
    bits as the default launch: same MMAs per tile, only the operand delivery differs.  The switch is
    read once per process, hence the child process."""
    import os
    import subprocess
    import sys
    code = ("import torch, hashlib\n"
            "from nesie_b200.linear_rows import gemm_nt\n"
            "torch.manual_seed(0)\n"
            "out = []\n"
            "for R, N, K in ((65536, 128, 256), (70000 - 64, 256, 132)):\n"
            "    a = torch.randn(R, K, device='cuda'); w = torch.randn(N, K, device='cuda')\n"
            "    out.append(hashlib.sha256(gemm_nt(a, w).cpu().numpy().tobytes()).hexdigest())\n"
            "print(' '.join(out))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for flag in ("0", "1"):
        env = dict(os.environ, NESIE_GEMM_PAIR=flag, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(r.stdout.strip().splitlines()[-1])
    assert res[0] == res[1]


def test_wgrad_wider_than_one_accumulator():
    """gy with more than 256 columns (VoteModule's 259-channel output conv): column blocks, no library GEMM."""
    from nesie_b200.linear_rows import wgrad
    torch.manual_seed(4)
    gy = torch.randn(8192, 259, device="cuda")
    x = torch.randn(8192, 256, device="cuda")
    want = gy.double().t() @ x.double()
    got = wgrad(gy, x)
    assert got.shape == (259, 256)
    assert ((got.double() - want).abs().max() / want.abs().max()).item() < 6e-6
