"""Dev tool: how close every gradient check of the parity tests sits to the tolerance (error / tolerance per parameter),
for the train-compare cases of tests/test_gpu_scale.py and tests/test_gpu_model.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import data_all, full_batch, make_graph, grad_close
from tests.test_gpu_model import cuda_net
from dgnn_b200 import runModel as rm
from oracle import trainer as otr
from oracle.static_model import SurfaceNet as OracleNet, make_clf

def run(tag, kw, npts, gseed, seed):
    g = make_graph(npts, seed=gseed); d = data_all(g, with_pos=True); data = full_batch(d)
    torch.manual_seed(seed)
    ref = OracleNet(make_clf(**kw))
    with torch.no_grad():
        for k, p in ref.named_parameters():
            if "norm" in k or k.startswith("decoder.1"):
                p.add_(0.3 * torch.randn_like(p))
    net = cuda_net(kw, ref.state_dict()); ref.train(); net.train()
    clf = make_clf(device="cuda:0", **kw)
    zr = ref(data); lr, _, _ = otr.cell_loss(zr, d.y, d.x[:, 0], clf.training.loss, clf.regularization.cell_norm, clf.regularization.cell_type)
    lr.backward()
    z = net(data); rm.cell_loss(z, d.y, d.x, clf).backward()
    refp = dict(ref.named_parameters())
    rows = sorted(((grad_close(p.grad, refp[k].grad)[0], k) for k, p in net.named_parameters()), reverse=True)
    print("%-10s worst: %s" % (tag, ", ".join("%s %.2f" % (k, e) for e, k in rows[:4])))

run("kf96-1500", {}, 1500, 61, 2)
run("eth", dict(convs=(64, 128, 256, 512)), 1500, 61, 2)
run("modelnet", dict(convs=(128, 256, 512, 1024)), 1500, 61, 2)
for s in (3, 4, 5):
    run("modelnet s%d" % s, dict(convs=(128, 256, 512, 1024)), 1500, 61 + s, 2 + s)
    run("kf96 s%d" % s, {}, 1500, 61 + s, 2 + s)
