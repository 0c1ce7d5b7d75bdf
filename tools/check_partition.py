"""Multi-GPU check (run under torchrun): partitioned inference vs the single-GPU forward, and its throughput.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/check_partition.py [n_points | nx lattice]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dgnn_b200 import synthetic as syn
from dgnn_b200.partition import PartitionedInference
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = "cuda:%d" % lr
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
if len(sys.argv) > 2 and sys.argv[2] == "lattice":      # analytic 4-regular scene with 2 * npts^3 cells (no qhull)
    adj, infinite, cen = syn.lattice_graph(npts, npts, npts)
else:
    pts = syn.random_points(npts, seed=0)
    adj, infinite, cen, _ = syn.delaunay_graph(pts)
n = infinite.shape[0]
x, ea, y = syn.synthetic_features(n, infinite, seed=1)
d = syn.to_attr(dict(x=torch.from_numpy(x), edge_attr=torch.from_numpy(ea), edge_index=torch.from_numpy(adj.T.astype(np.int64)).contiguous(),
                     pos=torch.from_numpy(cen.astype(np.float32))))
clf = syn.make_clf(device=dev)
torch.manual_seed(0)
net = SurfaceNet(clf).to(dev).eval()
with torch.no_grad():
    for blk in net.convs:                      # non-trivial running statistics
        blk.norm.module.running_mean.normal_(0, 0.1); blk.norm.module.running_var.uniform_(0.5, 1.5)
if world > 1:
    for t in list(net.parameters()) + list(net.buffers()):
        if t.is_floating_point():
            dist.broadcast(t.data, 0)
ref = net.inference_layer(d)
pi = PartitionedInference(net)
ids, out = pi.run(d)
err = (out - ref[ids]).abs().max().item()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): pi.run(d)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0.record()
for _ in range(10): pi.run(d)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
t = torch.tensor([ms, err], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
g, maps = pi._plan[0], pi._plan[1]
print("rank %d/%d: n=%d own=%d halo=%d (%.1f%%) max|partitioned - single| = %.2e" % (rank, world, n, maps.n_own, maps.n_halo, 100.0 * maps.n_halo / max(maps.n_own, 1), err), flush=True)
if rank == 0:
    print("PARTITIONED_INFERENCE world=%d cells=%d ms=%.3f cells/s=%.3e max_err=%.2e" % (world, n, t[0].item(), n / (t[0].item() * 1e-3), t[1].item()), flush=True)
assert t[1].item() < 1e-5
if world > 1: dist.destroy_process_group()
