"""Multi-GPU check (run under torchrun): one training step of a scene partitioned over the ranks
(halo exchange forward + backward, global BatchNorm statistics, global loss normaliser, gradient all-reduce)
against the same step on one GPU; then its throughput.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/check_partition_train.py [n_points]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dgnn_b200 import synthetic as syn
from dgnn_b200.partition import PartitionedTraining
from dgnn_b200.runModel import cell_loss
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = "cuda:%d" % lr
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
pts = syn.random_points(npts, seed=0)
adj, infinite, cen, _ = syn.delaunay_graph(pts)
n = infinite.shape[0]
x, ea, y = syn.synthetic_features(n, infinite, seed=1)
d = syn.to_attr(dict(x=torch.from_numpy(x), edge_attr=torch.from_numpy(ea), y=torch.from_numpy(y),
                     edge_index=torch.from_numpy(adj.T.astype(np.int64)).contiguous(),
                     pos=torch.from_numpy(cen.astype(np.float32))))
clf = syn.make_clf(device=dev)


def fresh():
    torch.manual_seed(0)
    return SurfaceNet(clf).to(dev).train()


# ---- single-GPU step on the whole scene (every rank computes it; identical weights by the seed)
ref = fresh()
ei = d.edge_index
batch = syn.to_attr(dict(all=d, batch_n_id=torch.arange(n), batch_adjs=[(ei, torch.arange(ei.shape[1]), (n, n))] * 5))
z_ref = ref(batch)
loss_ref = cell_loss(z_ref, d.y, d.x, clf)
loss_ref.backward()

# ---- the same step, partitioned
net = fresh()
pt = PartitionedTraining(net)
ids, z = pt.forward(d)
loss = pt.loss(z, d)
loss.backward()
pt.allreduce_gradients()
torch.cuda.synchronize()

err_z = ((z.detach() - z_ref.detach()[ids]).abs().max() / z_ref.detach().abs().mean()).item()
err_l = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
worst, worst_name = 0.0, ""
for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
    nr = q.grad.norm().item()
    e = (p.grad - q.grad).norm().item() / nr if nr > 1e-6 else (p.grad - q.grad).abs().max().item()
    if e > worst:
        worst, worst_name = e, k
for (k, b), (_, c) in zip(net.named_buffers(), ref.named_buffers()):
    if b.is_floating_point():
        e = ((b - c).abs().max() / (c.abs().max() + 1e-12)).item()
        assert e < 1e-4, ("running statistic", k, e)
g, maps = pt._plan[0], pt._plan[1]
print("rank %d/%d: own=%d halo=%d  logits rel err %.2e  loss rel err %.2e  worst grad rel err %.2e (%s)"
      % (rank, world, maps.n_own, maps.n_halo, err_z, err_l, worst, worst_name), flush=True)
assert err_z < 1e-4 and err_l < 1e-5 and worst < 2e-2, (err_z, err_l, worst, worst_name)

# ---- throughput of the partitioned step (fwd + loss + bwd + gradient all-reduce)
def step():
    for p in net.parameters():
        p.grad = None
    _, zz = pt.forward(d)
    pt.loss(zz, d).backward()
    pt.allreduce_gradients()

for _ in range(3): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("PARTITIONED_TRAINING world=%d cells=%d ms=%.3f cells/s=%.3e" % (world, n, t[0].item(), n / (t[0].item() * 1e-3)), flush=True)
if world > 1: dist.destroy_process_group()
