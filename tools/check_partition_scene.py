"""Multi-GPU check (run under torchrun): sharded build of a scene (each rank holds only its shard; halo maps negotiated
between the ranks; boundary-first order with the exchange overlapped) against the single-GPU path - inference logits
and one training step (loss + every gradient).  Two scenes: a real Delaunay graph cut into shards with
scene_from_global, and the analytic lattice scene of the benchmark.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/check_partition_scene.py [n_points]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dgnn_b200 import runModel as rm, scene as sc, synthetic as syn
from dgnn_b200.partition import PartitionedInference, PartitionedTraining
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = "cuda:%d" % lr
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 8000


def bcast(net):
    if world > 1:
        for t in list(net.parameters()) + list(net.buffers()):
            if t.is_floating_point():
                dist.broadcast(t.data, 0)


def gather_rows(ids, out, n):
    full = torch.zeros((n, out.shape[1]), dtype=torch.float32, device=dev)
    full[ids] = out
    if world > 1:
        dist.all_reduce(full)
    return full


def check(name, d, shard_fn):
    n = d.x.shape[0]
    clf = syn.make_clf(device=dev)
    torch.manual_seed(0)
    net = SurfaceNet(clf).to(dev)
    with torch.no_grad():
        for blk in net.convs:
            blk.norm.module.running_mean.normal_(0, 0.1); blk.norm.module.running_var.uniform_(0.5, 1.5)
    bcast(net)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    # inference
    net.eval()
    ref = net.inference_layer(d)
    for overlap in (True, False):
        pi = PartitionedInference(net)
        _, maps, _, ids, _ = pi.prepare_scene(shard_fn(False), overlap=overlap)
        ids, out = pi.run()
        err = (gather_rows(ids, out, n) - ref).abs().max().item()
        print("%s rank %d/%d overlap=%d own=%d halo=%d boundary=%d inference max|partitioned - single| = %.2e" % (
            name, rank, world, overlap, maps.n_own, maps.n_halo, maps.n_boundary, err), flush=True)
        assert err < 1e-5, err
    # one training step
    net.train()
    pt = PartitionedTraining(net)
    pt.prepare_scene(shard_fn(True))
    ids, logits = pt.forward()
    loss = pt.loss(logits)
    loss.backward()
    pt.allreduce_gradients()
    ref_net = SurfaceNet(clf).to(dev).train()
    ref_net.load_state_dict(state)
    ei = d.edge_index
    batch = syn.to_attr(dict(all=d, batch_n_id=torch.arange(n), batch_adjs=[(ei, torch.arange(ei.shape[1]), (n, n))] * 5))
    z = ref_net(batch)
    lref = rm.cell_loss(z, d.y, d.x, clf)
    lref.backward()
    zerr = (gather_rows(ids, logits.detach(), n) - z.detach()).abs().max().item()
    worst = 0.0
    refp = dict(ref_net.named_parameters())
    for k, p in net.named_parameters():
        g = refp[k].grad
        if float(g.norm()) < 1e-6:
            continue
        worst = max(worst, float((p.grad - g).norm() / g.norm()))
    rbuf = dict(ref_net.named_buffers())
    for k, b in net.named_buffers():
        if b.is_floating_point():
            assert torch.allclose(b, rbuf[k], rtol=1e-4, atol=1e-6), k
    print("%s rank %d/%d train: logits %.2e loss %.8f vs %.8f worst grad rel %.2e" % (name, rank, world, zerr, loss.item(),
                                                                                   lref.item(), worst), flush=True)
    assert zerr < 1e-4 and abs(loss.item() - lref.item()) < 2e-6 * max(1.0, abs(lref.item())) and worst < 2e-3, (zerr, worst)


pts = syn.random_points(npts, seed=0)
adj, infinite, cen, _ = syn.delaunay_graph(pts)
n = infinite.shape[0]
x, ea, y = syn.synthetic_features(n, infinite, seed=1)
d = syn.to_attr(dict(x=torch.from_numpy(x), edge_attr=torch.from_numpy(ea), y=torch.from_numpy(y),
                     edge_index=torch.from_numpy(adj.T.astype(np.int64)).contiguous(), pos=torch.from_numpy(cen.astype(np.float32))))
check("delaunay", d, lambda bwd: sc.scene_from_global(d, rank, world, dev, need_backward=bwd))
dims = (16, 16, 32)
gl = syn.to_attr(sc.lattice_global(dims))
check("lattice", gl, lambda bwd: sc.lattice_scene(dims, rank, world, dev, need_backward=bwd))
if rank == 0:
    print("PARTITIONED_SCENE world=%d ok" % world, flush=True)
if world > 1:
    dist.destroy_process_group()
