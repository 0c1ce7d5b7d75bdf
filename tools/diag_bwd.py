"""Dev tool: backward intermediates of the CUDA path against an fp64 restatement with autograd hooks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from dgnn_b200 import synthetic as og, runModel as rm, engine
from dgnn_b200.synthetic import make_clf, to_attr
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
from oracle import trainer as otr
from oracle.static_model import SurfaceNet as OracleNet

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 400
CONVS = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (64, 128, 128, 128)
pts = og.scan_like_points(npts, seed=0)
adj, infinite, cen, _ = og.delaunay_graph(pts)
n = infinite.shape[0]
x, ea, y = og.synthetic_features(n, infinite, seed=1)
ei = torch.from_numpy(adj.T.astype(np.int64)).contiguous()
torch.manual_seed(0)
ref32 = OracleNet(make_clf(convs=CONVS))
ref = OracleNet(make_clf(convs=CONVS)).double(); ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in ref32.state_dict().items()})
ref.train()
X = torch.from_numpy(x).double(); EA = torch.from_numpy(ea).double(); Y = torch.from_numpy(y).double()
h = X[:, 1:]
inter = {}
for l, blk in enumerate(ref.convs):
    z = blk.conv((h, h), EA, ei); z.retain_grad(); inter["z_%d" % l] = z
    yb = blk.norm(z); yb.retain_grad(); inter["y_%d" % l] = yb
    h = F.relu(yb)
zd = ref.decoder[0](h); zd.retain_grad(); inter["z_d"] = zd
yd = ref.decoder[1](zd); yd.retain_grad(); inter["y_d"] = yd
logits = ref.decoder[3](F.relu(yd)); logits.retain_grad()
lr, _, _ = otr.cell_loss(logits, Y, X[:, 0]); lr.backward()

clf = make_clf(device="cuda:0", convs=CONVS)
net = SurfaceNet(clf); net.load_state_dict(ref32.state_dict()); net.to("cuda:0").train()
d32 = to_attr(dict(x=torch.from_numpy(x), edge_attr=torch.from_numpy(ea), y=torch.from_numpy(y), edge_index=ei))
data32 = to_attr(dict(all=d32, batch_n_id=torch.arange(n), batch_adjs=[(ei, torch.arange(ei.shape[1]), (n, n))] * 5))
net.cache_graphs = True
engine.DEBUG = {}
zz = net(data32); loss = rm.cell_loss(zz, d32.y, d32.x, clf)
zz.retain_grad()
loss.backward()
g = data32._dgnn_plan[1][0]
perm = g.perm.long().cpu() if g.perm is not None else torch.arange(n)
def rel(a, b):
    rms = (b.norm() / b.numel() ** 0.5).item()
    return "frob %.2e med %.2e max %.2e" % (((a - b).norm() / b.norm()).item(), ((a - b).abs().median() / rms).item(), ((a - b).abs().max() / rms).item())
print("dlogits rel", rel(zz.grad.cpu().double(), logits.grad))
D = engine.DEBUG
# dy (masked grad at norm output) in internal order -> compare with oracle y.grad * (y>0)
def dy_ref(name):
    yb = inter[name]
    return (yb.grad * (yb > 0))[perm]
print("dy_d rel", rel(D["dy_d"].cpu().double(), dy_ref("y_d")))
for l in range(3, -1, -1):
    print("dy_%d rel" % l, rel(D["dy_%d" % l].cpu().double(), dy_ref("y_%d" % l)))
print("dh_L rel (masked, in-place)", rel(D["dh_L"].cpu().double(), dy_ref("y_3")))
