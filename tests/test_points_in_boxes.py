"""points_in_boxes: C oracle vs outputs recorded from the reference kernels
(tests/golden/pib_golden.npz), and on the GPU this repo's kernels vs the oracle, the golden vectors
and the reference kernels live."""
import os

import numpy as np
import pytest
import torch

from oracle import cpu, ref_cuda

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pib_golden.npz")


def _cases():
    g = np.load(GOLD)
    for i in range(int(g["count"])):
        yield (i, torch.from_numpy(g[f"c{i}_pts"]), torch.from_numpy(g[f"c{i}_boxes"]),
               torch.from_numpy(g[f"c{i}_first"]), torch.from_numpy(g[f"c{i}_batch"].astype(np.int32)),
               bool(g[f"c{i}_yaw"]))


def test_oracle_matches_reference_kernel_outputs():
    for i, pts, boxes, first, batch, yaw in _cases():
        got_b = cpu.points_in_boxes_batch(pts, boxes)
        got_f = cpu.points_in_boxes_gpu(pts, boxes)
        if not yaw:  # the yaw-0 boxes of this path: bit-exact, face points included
            assert torch.equal(got_b, batch), i
            assert torch.equal(got_f, first), i
        else:        # libm vs CUDA sinf / cosf: only points within rounding of a face may differ
            assert (got_b != batch).float().mean() < 1e-5, i
        assert batch.sum() > 0


def test_wrapper_contract_on_cpu_tensors():
    import nesie_b200 as nb
    with pytest.raises((RuntimeError, AssertionError)):
        nb.points_in_boxes_batch(torch.zeros(1, 4, 3), torch.zeros(1, 2, 7))


@pytest.mark.gpu
def test_gpu_matches_golden_and_oracle():
    import nesie_b200 as nb
    for i, pts, boxes, first, batch, yaw in _cases():
        got_b = nb.points_in_boxes_batch(pts.cuda(), boxes.cuda())
        got_f = nb.points_in_boxes_gpu(pts.cuda(), boxes.cuda())
        assert got_b.dtype == torch.int32 and got_f.dtype == torch.int32
        assert torch.equal(got_b.cpu(), batch), i          # same arithmetic as the reference: every yaw
        assert torch.equal(got_f.cpu(), first), i
        if not yaw:
            assert torch.equal(got_b.cpu(), cpu.points_in_boxes_batch(pts, boxes))


@pytest.mark.gpu
def test_gpu_matches_reference_kernel_live_at_scene_size():
    import nesie_b200 as nb
    if not ref_cuda.pib_available():
        pytest.skip("oracle/_ref/libnesie_ref_pib.so not built")
    g = torch.Generator().manual_seed(9)
    B, M, T = 8, 40000, 64
    pts = (torch.rand(B, M, 3, generator=g) * 8 - 4).cuda()
    ctr = torch.rand(B, T, 3, generator=g) * 6 - 3
    boxes = torch.cat([ctr, torch.rand(B, T, 3, generator=g) * 2 + 0.2,
                       (torch.rand(B, T, 1, generator=g) - 0.5) * 6], -1).cuda()
    assert torch.equal(nb.points_in_boxes_batch(pts, boxes), ref_cuda.points_in_boxes_batch(pts, boxes))
    assert torch.equal(nb.points_in_boxes_gpu(pts, boxes), ref_cuda.points_in_boxes_gpu(pts, boxes))
    # more boxes than one staging pass, ragged sizes
    T2 = 1100
    boxes2 = torch.cat([torch.rand(1, T2, 3, generator=g) * 6 - 3, torch.rand(1, T2, 3, generator=g) + 0.2,
                        torch.zeros(1, T2, 1)], -1).cuda()
    pts2 = (torch.rand(1, 777, 3, generator=g) * 8 - 4).cuda()
    assert torch.equal(nb.points_in_boxes_batch(pts2, boxes2), ref_cuda.points_in_boxes_batch(pts2, boxes2))
    assert torch.equal(nb.points_in_boxes_gpu(pts2, boxes2), ref_cuda.points_in_boxes_gpu(pts2, boxes2))
