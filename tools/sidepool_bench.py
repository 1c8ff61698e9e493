"""SidePooling quality head at the Nesie shape (8 scenes, 2 x 256 proposals, 1024 seeds x 256 ch, 18
classes): this repo's module vs the reference FORMULATION on the same GPU (torch gather / index_select
loop / cuDNN 1x1 convs, TF32 off; its three_nn is mmcv's, replaced here by this repo's kernel)."""
import copy, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nesie_b200 as nb
from nesie_b200.side_pooling import SidePooling
from oracle.side_pooling_ref import SidePoolingOracle

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


class RefStyle(SidePoolingOracle):
    """The oracle's reference-formulation hooks, on CUDA tensors."""

    def _grid_rows(self, origin_xyz, origin_features, grid, center):
        B, T = grid.shape[:2]
        K = center.shape[1]
        C = origin_features.shape[1]
        _, idx = nb.three_nn(grid, origin_xyz)
        idx = idx.long()
        near = torch.gather(origin_xyz, 1, idx.view(B, -1, 1).expand(-1, -1, 3))
        d = near - grid.unsqueeze(2).expand(-1, -1, 3, -1).reshape(B, -1, 3)
        dist = torch.sqrt(torch.sum(d * d, dim=2))
        weight = (1 / (dist + 1e-8)).view(B, -1, 3)
        weight = weight / torch.sum(weight, dim=2, keepdim=True)
        table = origin_features.transpose(1, 2)
        feats = torch.stack([table[b].index_select(0, idx[b].reshape(-1)) for b in range(B)], 0)
        feats = torch.sum(feats.view(B, -1, 3, C) * weight.unsqueeze(-1), dim=2)
        head = grid.view(B, K, T // K, 3) - center.unsqueeze(2)
        return torch.cat([head.reshape(B * T, 3), feats.reshape(B * T, C)], dim=1)


def inputs(B, K, N, C, dev):
    g = torch.Generator().manual_seed(3)
    center = torch.rand(B, K, 3, generator=g) * 6 - 3
    size = torch.rand(B, K, 3, generator=g) * 1.5 + 0.2
    heading = torch.zeros(B, K)
    ep = {"seed_points": torch.rand(B, N, 3, generator=g) * 8 - 4,
          "seed_features": torch.randn(B, C, N, generator=g),
          "bbox_probs": torch.softmax(torch.randn(B, 6, 33, K // 2, generator=g), dim=2)}
    return [t.to(dev) for t in (center, size, heading)], {k: v.to(dev) for k, v in ep.items()}


def timeit(mod, boxes, ep, train, iters=5):
    def step():
        out = mod(*boxes, dict(ep))
        if train:
            (out["side_scores"].sum() + out["iou_scores"].sum()).backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


B, K, N, C, NC = 8, 512, 1024, 256, 18
torch.manual_seed(0)
base = SidePoolingOracle(NC, 1, NC, None, K // 2, "vote", seed_feat_dim=C)
mine = SidePooling(NC, 1, NC, None, K // 2, "vote", seed_feat_dim=C)
mine.load_state_dict(copy.deepcopy(base.state_dict()))
refs = RefStyle(NC, 1, NC, None, K // 2, "vote", seed_feat_dim=C)
refs.load_state_dict(copy.deepcopy(base.state_dict()))
mine, refs = mine.cuda(), refs.cuda()
boxes, ep = inputs(B, K, N, C, "cuda")
rows = B * K * (96 + 64)
gflop = 2.0 * rows * ((C + 3) * 256 + 256 * 128 + 256 * 256 + 256 * 128) / 1e9
with torch.no_grad():
    a = mine(*boxes, dict(ep)); b = refs(*boxes, dict(ep))
err = max(float((a[k] - b[k]).abs().max() / b[k].abs().max()) for k in ("side_scores", "iou_scores"))
res = {"shape": f"B{B} K{K} N{N} C{C}", "grid_rows": rows, "mlp_gflop_fwd": round(gflop, 1), "rel_err_vs_ref_style": err}
for train in (False, True):
    mine.train(train); refs.train(train)
    ctx = torch.enable_grad() if train else torch.no_grad()
    with ctx:
        t1 = timeit(mine, boxes, ep, train); t2 = timeit(refs, boxes, ep, train)
    res["train_fwd_bwd" if train else "eval_fwd"] = {"this_repo_ms": round(t1, 3), "reference_style_ms": round(t2, 3),
                                                   "speed_up": round(t2 / t1, 2)}
# bounded CPU sample of the oracle (1 scene)
cb, ce = inputs(1, K, N, C, "cpu")
base.eval()
t0 = time.perf_counter()
with torch.no_grad():
    base(*cb, dict(ce))
res["cpu_oracle_eval_fwd_1_scene_s"] = round(time.perf_counter() - t0, 2)
print(json.dumps(res))
