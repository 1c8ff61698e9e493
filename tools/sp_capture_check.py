import sys, torch, traceback
sys.path.insert(0, "/root/repo")
from nesie_b200.side_pooling import SidePooling
torch.manual_seed(0)
B,K,N,C,NC=2,64,256,256,18
m = SidePooling(NC,1,NC,None,K//2,"vote",seed_feat_dim=C).cuda()
g = torch.Generator().manual_seed(1)
center=(torch.rand(B,K,3,generator=g)*4-2).cuda(); size=(torch.rand(B,K,3,generator=g)+0.2).cuda(); heading=torch.zeros(B,K).cuda()
ep={"seed_points":(torch.rand(B,N,3,generator=g)*5-2.5).cuda(),"seed_features":torch.randn(B,C,N,generator=g).cuda(),"bbox_probs":torch.softmax(torch.randn(B,6,33,K//2,generator=g),2).cuda()}
def f():
    out = m(center,size,heading,dict(ep))
    (out["side_scores"].sum()+out["iou_scores"].sum()).backward()
s=torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): f()
torch.cuda.synchronize()
gr=torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(gr):
        f()
    gr.replay(); torch.cuda.synchronize(); print("capture ok")
except Exception:
    traceback.print_exc()
