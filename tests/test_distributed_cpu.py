"""N > 1 path on the CPU (gloo, world_size 2) through the PRODUCT's data-parallel class
(nesie_b200.ddp.FlatGradDDP: flat gradient buffer, bucketed all-reduce launched from
post-accumulate hooks): scenes shard by rank, replicas start identical, the exchanged gradient is the
mean of the two ranks' gradients bit for bit, and both ranks end the optimizer step with identical
weights.  The model is the CPU twin of the VoteNet step (the product's kernels need a GPU)."""
import os
import socket
import sys

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


def tiny_votenet(cls):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from head_cases import mean_size_file
    from nesie_b200.detectors import nesie_head_cfg
    head = nesie_head_cfg(mean_size_arr_path=mean_size_file(), num_proposal=16)
    head["vote_module_cfg"] = dict(head["vote_module_cfg"], in_channels=32, conv_channels=(32, 32))
    head["vote_aggregation_cfg"] = dict(head["vote_aggregation_cfg"], mlp_channels=[32, 32, 32, 32], num_sample=8)
    head["pred_layer_cfg"] = dict(in_channels=32, shared_conv_channels=(32, 32), bias=True)
    head["grid_conv_cfg"] = dict(head["grid_conv_cfg"], seed_feat_dim=32)
    backbone = dict(in_channels=4, num_points=(128, 64, 32, 16), radius=(0.4, 0.8, 1.2, 1.6),
                    num_samples=(8, 8, 8, 8), sa_channels=((16, 16, 32), (32, 32, 32), (32, 32, 32), (32, 32, 32)),
                    fp_channels=((32, 32), (32, 32)))
    return cls(backbone=backbone, bbox_head=head)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nesie_b200.ddp import FlatGradDDP
    from nesie_b200.synthetic import make_batch
    from oracle.detectors_ref import VoteNetRef
    torch.set_num_threads(2)
    torch.manual_seed(rank)                       # replicas start DIFFERENT: broadcast must fix it
    model = tiny_votenet(VoteNetRef)
    ddp = FlatGradDDP(model, bucket_bytes=64 << 10)
    assert len(ddp.buckets) > 2
    opt = torch.optim.AdamW(ddp.params, lr=0.008, weight_decay=0.01)
    pts, gb, gl = make_batch(2, 1024, seed0=100 * rank, origin='bottom')   # each rank: its own scenes
    ddp.zero_grad()
    torch.manual_seed(7)                          # same jitter noise stream on both ranks
    loss = sum(model.forward_train(pts, gb, gl).values())
    loss.backward()
    ddp.finish()
    reduced = ddp.flat.clone()
    # the local (pre-exchange) gradient is gone once the hooks have fired: recompute it without them
    ddp.remove_hooks()
    ddp.zero_grad()
    torch.manual_seed(7)
    sum(model.forward_train(pts, gb, gl).values()).backward()
    for p in ddp.params:                           # without hooks the gradients are plain tensors
        if p.grad is not None:
            ddp._view[id(p)].copy_(p.grad)
    local = ddp.flat.clone()
    both = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(both, local)
    ddp.flat.copy_(reduced)
    for p in ddp.params:
        p.grad = ddp._view[id(p)]
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
        mean = (both[0] + both[1]) / world
        out["grad_is_mean"] = bool(torch.allclose(reduced, mean, rtol=1e-6, atol=1e-9))
        out["grad_nonzero"] = bool(reduced.abs().sum() > 0)
    dist.destroy_process_group()


def test_two_rank_gloo_step_keeps_replicas_identical():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["finite"] and out["same_weights"] and out["different_data"]
    assert out["grad_is_mean"] and out["grad_nonzero"]
