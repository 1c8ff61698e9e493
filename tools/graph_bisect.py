import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesie_b200 as nb
from nesie_b200.synthetic import make_batch
from nesie_b200.votenet import VoteNetHarness
torch.manual_seed(0)
model = VoteNetHarness().cuda()
opt = torch.optim.AdamW(model.parameters(), lr=0.008, weight_decay=0.01, fused=True, capturable=True)
pts, gb, gl = make_batch(8, 40000)
pts = pts.cuda(); gb = [b.cuda() for b in gb]; gl = [l.cuda() for l in gl]
xyz = pts[..., :3].contiguous()

def try_capture(name, fn):
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "OK", flush=True)
    except Exception as e:
        print(name, "FAILED", repr(e)[:160].replace("\n", " "), flush=True)
        torch.cuda.synchronize()

def nograd(f):
    def g():
        with torch.no_grad():
            return f()
    return g
try_capture("fps_small(2048->1024)", nograd(lambda: nb.furthest_point_sample(xyz[:, :2048].contiguous(), 1024)))
try_capture("fps_cluster(40000->2048)", nograd(lambda: nb.furthest_point_sample(xyz, 2048)))
c = xyz[:, :2048].contiguous()
try_capture("ball_query_grid", nograd(lambda: nb.ball_query(0.0, 0.2, 64, xyz, c)))
try_capture("ball_query_brute", nograd(lambda: nb.ball_query(0.0, 0.4, 32, c, c[:, :1024].contiguous())))
model.backbone.overlap_fps = False
try_capture("backbone_fwd_no_overlap", nograd(lambda: model.backbone(pts)))
model.backbone.overlap_fps = True
try_capture("backbone_fwd_overlap", nograd(lambda: model.backbone(pts)))
try_capture("forward", nograd(lambda: model(pts)))
try_capture("loss fwd only", nograd(lambda: model.train_step_loss(pts, gb, gl)))
f = torch.randn(8, 16, 2048, device="cuda", requires_grad=True)
idx = torch.randint(0, 2048, (8, 256, 16), dtype=torch.int32, device="cuda")
def op_bwd(fn):
    def g():
        f.grad = None
        fn().sum().backward()
    return g
try_capture("grouping bwd", op_bwd(lambda: nb.grouping_operation(f, idx)))
c2 = xyz[:, :256].contiguous()
grp = nb.QueryAndGroup(0.4, 16, normalize_xyz=True)
try_capture("query_and_group bwd", op_bwd(lambda: grp(xyz[:, :2048].contiguous(), c2, f)))
i3 = torch.randint(0, 2048, (8, 512, 3), dtype=torch.int32, device="cuda"); w3 = torch.rand(8, 512, 3, device="cuda")
try_capture("interpolate bwd", op_bwd(lambda: nb.three_interpolate(f, i3, w3)))
sa = nb.PointSAModule(mlp_channels=[16, 32, 32, 64], num_point=256, radius=0.4, num_sample=16, normalize_xyz=True).cuda()
try_capture("sa module bwd", op_bwd(lambda: sa(xyz[:, :2048].contiguous(), f)[1]))
def bb():
    model.zero_grad(set_to_none=False)
    model.backbone(pts)["fp_features"][-1].sum().backward()
try_capture("backbone fwd+bwd", bb)
def fwd_loss_bwd_nobackbone():
    model.zero_grad(set_to_none=False)
    loss, parts = model.train_step_loss(pts, gb, gl)
    parts["surface_loss"].backward()
try_capture("surface loss bwd", fwd_loss_bwd_nobackbone)
def fb():
    opt.zero_grad(set_to_none=False)
    loss, _ = model.train_step_loss(pts, gb, gl)
    loss.backward()
try_capture("fwd+bwd", fb)
def full():
    fb(); torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0); opt.step()
try_capture("full step", full)
