"""Dev tool: per-parameter gradient error of the CUDA path against the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dgnn_b200 import synthetic as og, runModel as rm
from dgnn_b200.synthetic import make_clf, to_attr
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
from oracle import trainer as otr
from oracle.static_model import SurfaceNet as OracleNet

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 400
pts = og.scan_like_points(npts, seed=0)
adj, infinite, cen, _ = og.delaunay_graph(pts)
n = infinite.shape[0]
x, ea, y = og.synthetic_features(n, infinite, seed=1)
ei = torch.from_numpy(adj.T.astype(np.int64)).contiguous()
def mk(dtype):
    d = to_attr(dict(x=torch.from_numpy(x).to(dtype), edge_attr=torch.from_numpy(ea).to(dtype), y=torch.from_numpy(y).to(dtype), edge_index=ei))
    return d, to_attr(dict(all=d, batch_n_id=torch.arange(n), batch_adjs=[(ei, torch.arange(ei.shape[1]), (n, n))] * 5))
torch.manual_seed(0)
ref32 = OracleNet(make_clf())
ref = OracleNet(make_clf()).double(); ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in ref32.state_dict().items()})
d64, data64 = mk(torch.float64)
ref.train(); zr = ref(data64); lr, _, _ = otr.cell_loss(zr, d64.y, d64.x[:, 0]); lr.backward()
clf = make_clf(device="cuda:0")
net = SurfaceNet(clf); net.load_state_dict(ref32.state_dict()); net.to("cuda:0").train()
d32, data32 = mk(torch.float32)
z = net(data32); loss = rm.cell_loss(z, d32.y, d32.x, clf); loss.backward()
print("n", n, "logit max err", (z.detach().cpu().double() - zr.detach()).abs().max().item(), "loss", loss.item(), lr.item())
refp = dict(ref.named_parameters())
for k, p in net.named_parameters():
    g = refp[k].grad; nr = g.norm().item()
    print("%-34s |g|=%.3e rel=%.3e" % (k, nr, ((p.grad.cpu().double() - g).norm() / max(nr, 1e-30)).item()))
