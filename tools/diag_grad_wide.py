"""Dev tool: per-parameter gradient error (relative Frobenius, median / RMS) of the CUDA path against the fp64 AND fp32
oracle for a width list:  python tools/diag_grad_wide.py 64,128,256,512 [n_points]"""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import data_all, full_batch, make_graph
from dgnn_b200 import runModel as rm
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
from oracle import trainer as otr
from oracle.static_model import SurfaceNet as OracleNet, make_clf, to_attr

convs = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (64, 128, 256, 512)
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
g = make_graph(npts, seed=61); d = data_all(g, with_pos=True)
torch.manual_seed(2)
ref = OracleNet(make_clf(convs=convs))
with torch.no_grad():
    for k, p in ref.named_parameters():
        if "norm" in k or k.startswith("decoder.1"):
            p.add_(0.3 * torch.randn_like(p))
ref64 = copy.deepcopy(ref).double()
ref.train(); ref64.train()
z32 = ref(full_batch(d)); otr.cell_loss(z32, d.y, d.x[:, 0])[0].backward()
d64 = to_attr(dict(x=d.x.double(), edge_attr=d.edge_attr.double(), y=d.y.double(), edge_index=d.edge_index))
z64 = ref64(full_batch(d64)); otr.cell_loss(z64, d64.y, d64.x[:, 0])[0].backward()
clf = make_clf(device="cuda:0", convs=convs)
net = SurfaceNet(clf); net.load_state_dict(ref.state_dict()); net.to("cuda:0").train()
z = net(full_batch(d)); rm.cell_loss(z, d.y, d.x, clf).backward()
print("logit err vs fp64: cuda %.2e  oracle32 %.2e" % ((z.detach().cpu().double() - z64.detach()).abs().max().item(),
                                                        (z32.detach().double() - z64.detach()).abs().max().item()))
p64, p32 = dict(ref64.named_parameters()), dict(ref.named_parameters())
def err(a, b):
    a = a.double().cpu(); b = b.double()
    nr = b.norm().item()
    if nr < 1e-12: return 0.0, 0.0
    rms = nr / b.numel() ** 0.5
    return ((a - b).norm() / nr).item(), ((a - b).abs().median() / rms).item()
for k, p in net.named_parameters():
    f1, m1 = err(p.grad, p64[k].grad); f2, m2 = err(p32[k].grad, p64[k].grad)
    print("%-34s cuda: frob %.2e med %.2e | oracle32: frob %.2e med %.2e" % (k, f1, m1, f2, m2))
