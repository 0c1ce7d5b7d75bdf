"""Dev: the bench's configs[1] training step (64 objects, kf96 widths) a few times, for ncu launch lists.
    python tools/step_once.py [n_objects] [steps] [infer]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dgnn_b200 import runModel as rm
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
from dgnn_b200.synthetic import make_clf, to_attr
nobj = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
infer = len(sys.argv) > 3 and sys.argv[3] == "infer"
dev = torch.device("cuda:0")
host = bench.make_objects(nobj, 0)
clf = make_clf(convs=bench.WIDTHS, device="cuda:0")
torch.manual_seed(0)
net = SurfaceNet(clf).to(dev).train()
opt = rm.Adam(net.parameters(), lr=0.005)
dres = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
data = bench.batch_of(dres, to_attr)
dall = to_attr({k: v for k, v in dres.items() if k != "n"})
print("cells", host["n"])
if infer:
    net.eval()
for i in range(steps):
    torch.cuda.synchronize()
    print("STEP", i, flush=True)
    if infer:
        with torch.no_grad():
            net.inference_layer(dall)
    else:
        loss = rm.cell_loss(net(data), data.all.y, data.all.x, clf)
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
torch.cuda.synchronize()
