"""Multi-GPU check (run under torchrun): one training step of the Updated-edge-filter model on a scene partitioned
over the ranks (BASELINE configs[3]) against the same step on one GPU; then its throughput.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/check_partition_upd.py [n_points] [sage|sage+]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dgnn_b200 import synthetic as syn
from dgnn_b200.partition import PartitionedUpdatedTraining
from dgnn_b200.surfaceNetUpdatedEdgeFilters import SurfaceNet

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = "cuda:%d" % lr
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
name = sys.argv[2] if len(sys.argv) > 2 else "sage+"
pts = syn.random_points(npts, seed=0)
adj, infinite, cen, _ = syn.delaunay_graph(pts)
n = infinite.shape[0]
x, ea, y = syn.synthetic_features(n, infinite, seed=1)
ei = torch.from_numpy(adj.T.astype(np.int64)).contiguous()
d = syn.to_attr(dict(x=torch.from_numpy(x), edge_attr=torch.from_numpy(ea), edge_index=ei,
                     pos=torch.from_numpy(cen.astype(np.float32))))
clf = syn.to_attr(dict(training=dict(model_params=[64, 128, 128, 128], model_name=name),
                       features=dict(normalization_feature=1, keep_normalization_feature=0), temp=dict(device=dev)))
w = torch.linspace(0.5, 1.5, n, device=dev)[:, None]


def fresh():
    torch.manual_seed(0)
    return SurfaceNet(28, clf).to(dev)


ref = fresh()
full = syn.to_attr(dict(x=d.x, edge_attr=d.edge_attr, n_id=torch.arange(n), adjs=[(ei, torch.arange(ei.shape[1]), (n, n))] * 4))
y_ref = ref(full)
(y_ref.square() * w).sum().backward()

net = fresh()
pt = PartitionedUpdatedTraining(net)
ids, out = pt.forward(d)
(out.square() * w[ids]).sum().backward()
pt.allreduce_gradients()
torch.cuda.synchronize()
err_y = ((out.detach() - y_ref.detach()[ids]).abs().max() / y_ref.detach().abs().mean()).item()
worst, worst_name = 0.0, ""
for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
    e = (p.grad - q.grad).norm().item() / max(q.grad.norm().item(), 1e-12)
    if e > worst:
        worst, worst_name = e, k
maps = pt._plan[1]
print("rank %d/%d: own=%d halo=%d  output rel err %.2e  worst grad rel err %.2e (%s)"
      % (rank, world, maps.n_own, maps.n_halo, err_y, worst, worst_name), flush=True)
assert err_y < 1e-4 and worst < 2e-2, (err_y, worst, worst_name)


def step():
    for p in net.parameters():
        p.grad = None
    ids, o = pt.forward(d)
    (o.square() * w[ids]).sum().backward()
    pt.allreduce_gradients()


for _ in range(3): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("PARTITIONED_UPDATED_TRAINING world=%d model=%s cells=%d ms=%.3f cells/s=%.3e" % (world, name, n, t[0].item(), n / (t[0].item() * 1e-3)), flush=True)
if world > 1: dist.destroy_process_group()
