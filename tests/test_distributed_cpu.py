"""N > 1 path on the CPU (gloo, world_size 2): scenes shard by rank, DDP averages gradients, and
every rank ends a step with identical weights -- the only collective on this path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nesie_b200.synthetic import make_batch
    from oracle.votenet_ref import VoteNetOracle
    torch.manual_seed(0)
    torch.set_num_threads(2)
    model = VoteNetOracle(num_points=(128, 64, 32, 16), num_samples=(8, 8, 8, 8), num_proposal=16)

    class Step(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, pts, gb, gl):
            return self.m.train_step_loss(pts, gb, gl)[0]

    ddp = torch.nn.parallel.DistributedDataParallel(Step(model), broadcast_buffers=False)
    opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01)
    pts, gb, gl = make_batch(2, 2048, seed0=100 * rank)  # each rank: its own scenes
    loss = ddp(pts, gb, gl)
    loss.backward()
    opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    losses = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(losses, loss.detach().reshape(1))
    if rank == 0:
        out["same_weights"] = bool(torch.equal(gathered[0], gathered[1]))
        out["different_data"] = bool(abs(float(losses[0]) - float(losses[1])) > 0)
        out["finite"] = bool(torch.isfinite(flat).all())
    dist.destroy_process_group()


def test_two_rank_gloo_step_keeps_replicas_identical():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["finite"] and out["same_weights"] and out["different_data"]
