"""Dev: per C-ABI call time of the Updated-edge-filter training step of bench.updated_training (world 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from dgnn_b200 import runModel as rm, scene as sc
from dgnn_b200.partition import PartitionedUpdatedTraining
from dgnn_b200.surfaceNetUpdatedEdgeFilters import SurfaceNet as UpdNet
from dgnn_b200.synthetic import make_clf, to_attr
dev = torch.device("cuda:0")
dims = bench.UPD_DIMS[1]
clf = to_attr(dict(training=dict(model_params=list(bench.WIDTHS), model_name="sage+"),
                   features=dict(normalization_feature=0, keep_normalization_feature=0), temp=dict(device=str(dev))))
loss_clf = make_clf(device=str(dev))
torch.manual_seed(0)
net = UpdNet(bench.F0, clf).to(dev).train()
opt = rm.Adam(net.parameters(), lr=0.005)
shard = sc.lattice_scene(dims, 0, 1, dev)
pt = PartitionedUpdatedTraining(net); pt.prepare_scene(shard)
def step():
    _, logits = pt.forward()
    loss = rm.cell_loss(logits, shard.y, shard.w, loss_clf, group=None, distributed=False)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
print("step %.2f ms" % (e0.elapsed_time(e1) / 2))
prof = bench.KernelProfile(); prof.install()
step(); step()
prof.uninstall()
os.environ["DGNN_BENCH_TABLE"] = "40"
roof, table, total = prof.summary(2, 6451.8, "x")
print("C-ABI calls %.2f ms per step" % total)
for k in table: print(k)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    step(); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
