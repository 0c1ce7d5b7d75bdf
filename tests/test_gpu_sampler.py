"""On-device closure sampler against the oracle's restatement of the PyG NeighborSampler (bit-exact integer maps), and
a training step driven by it."""
import numpy as np
import pytest
import torch

from oracle.static_model import NeighborSampler as OracleSampler, SurfaceNet as OracleNet, make_clf, to_attr
from tests.helpers import data_all, grad_close, logits_close, make_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("hops,batch_size", [(1, 64), (5, 37), (5, 2048)])
def test_sampler_matches_oracle_bit_exact(hops, batch_size):
    from dgnn_b200.sampler import NeighborSampler
    g = make_graph(900, seed=12)
    d = data_all(g)
    n = d.x.shape[0]
    node_idx = torch.arange(3, n - 5, 2)
    ref = OracleSampler(d.edge_index, [-1] * hops, batch_size, node_idx=node_idx, num_nodes=n)
    smp = NeighborSampler(d.edge_index, [-1] * hops, batch_size=batch_size, node_idx=node_idx, num_nodes=n, device=DEV)
    assert len(smp) == len(ref)
    for (bs, n_id, adjs), (bs_r, n_id_r, adjs_r) in zip(smp, ref):
        assert bs == bs_r and torch.equal(n_id.cpu(), n_id_r)
        if hops == 1:
            adjs, adjs_r = [adjs], [adjs_r]
        assert len(adjs) == len(adjs_r)
        for (ei, e_id, size), (ei_r, e_id_r, size_r) in zip(adjs, adjs_r):
            assert tuple(size) == tuple(size_r)
            assert torch.equal(ei.cpu(), ei_r) and torch.equal(e_id.cpu(), e_id_r)
    # the scratch tables are back to idle: a second pass gives the same batches
    first = next(iter(smp))
    again = next(iter(smp))
    assert torch.equal(first[1], again[1])


def test_training_step_driven_by_device_sampler():
    """Seed-batch training as in run.py -t: the sampled closure goes straight from the device sampler into the model;
    logits and gradients match the oracle model fed by the oracle sampler."""
    from dgnn_b200 import runModel as rm
    from dgnn_b200.sampler import NeighborSampler
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from oracle import trainer as otr
    g = make_graph(700, seed=13)
    d = data_all(g)
    n = d.x.shape[0]
    kw = dict(convs=(16, 32, 32, 32))
    torch.manual_seed(1)
    ref = OracleNet(make_clf(**kw)).train()
    clf = make_clf(device=DEV, **kw)
    net = SurfaceNet(clf)
    net.load_state_dict(ref.state_dict(), strict=True)
    net.to(DEV).train()
    seeds = torch.arange(50, 178)
    bs, n_id_r, adjs_r = next(iter(OracleSampler(d.edge_index, [-1] * 5, 128, node_idx=seeds, num_nodes=n)))
    bs2, n_id, adjs = next(iter(NeighborSampler(d.edge_index, [-1] * 5, batch_size=128, node_idx=seeds, num_nodes=n, device=DEV)))
    zr = ref(to_attr(dict(all=d, batch_n_id=n_id_r, batch_adjs=adjs_r)))
    z = net(to_attr(dict(all=d, batch_n_id=n_id, batch_adjs=adjs)))
    err, ok = logits_close(z.detach().cpu().numpy(), zr.detach().numpy())
    assert ok, err
    n_sup = z.shape[0]
    sup = n_id_r[:n_sup]
    lr, _, _ = otr.cell_loss(zr, d.y[sup], d.x[sup][:, 0], "kl", None, "vol")
    lr.backward()
    loss = rm.cell_loss(z, d.y[sup], d.x[sup], clf)
    loss.backward()
    assert abs(loss.item() - lr.item()) <= 2e-5 * max(1.0, abs(lr.item()))
    refp = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        e, tol = grad_close(p.grad, refp[k].grad)
        assert e <= tol, (k, e, tol)


def test_sampler_bool_mask_drop_last_and_adj_interface():
    """run.py:51,72-74 passes reduceDataset's BOOL train_mask with drop_last=True, sampler=None, return_e_id=edge_convs and
    reads adj.size / adj.edge_index (runModel.py:119,274)."""
    from dgnn_b200.sampler import Adj, NeighborSampler
    g = make_graph(500, seed=21)
    d = data_all(g)
    n = d.x.shape[0]
    rng = np.random.default_rng(5)
    mask = torch.from_numpy(rng.random(n) < 0.4)
    ref = OracleSampler(d.edge_index, [-1] * 3, 100, node_idx=mask, drop_last=True)
    smp = NeighborSampler(edge_index=d.edge_index, node_idx=mask, sizes=[-1] * 3, batch_size=100, sampler=None, shuffle=False,
                          drop_last=True, return_e_id=1, device=DEV)
    assert len(smp) == len(ref) == int(mask.sum()) // 100
    batches = list(smp)
    assert len(batches) == len(ref)
    for (bs, n_id, adjs), (bs_r, n_id_r, adjs_r) in zip(batches, ref):
        assert bs == bs_r == 100 and torch.equal(n_id.cpu(), n_id_r)
        for a, (ei_r, e_id_r, size_r) in zip(adjs, adjs_r):
            assert isinstance(a, Adj) and a.size[1] == size_r[1] and tuple(a.size) == tuple(size_r)
            assert torch.equal(a.edge_index.cpu(), ei_r) and torch.equal(a.e_id.cpu(), e_id_r)
            assert a.to("cpu").edge_index.device.type == "cpu"
    # seeds are the mask's nonzero positions, not the 0/1 values of the mask
    assert torch.equal(torch.cat([b[1][:b[0]] for b in batches]).cpu(), mask.nonzero().view(-1)[:len(ref) * 100])
    # return_e_id = 0 (edge_convs 0): no e_id
    smp0 = NeighborSampler(d.edge_index, [-1], batch_size=64, return_e_id=0, device=DEV)
    assert next(iter(smp0))[2].e_id is None
    with pytest.raises(ValueError):
        NeighborSampler(d.edge_index, [-1], batch_size=4, node_idx=torch.tensor([1, 1, 2]), num_nodes=n, device=DEV)
    with pytest.raises(ValueError):
        NeighborSampler(d.edge_index, [-1], batch_size=4, node_idx=torch.tensor([1, n]), num_nodes=n, device=DEV)


def test_sampler_orders_in_edges_by_source_on_shuffled_edge_lists():
    """A general edge list (not the 4i+k file layout): PyG's adj_t rows are sorted by source id, not by edge id."""
    from dgnn_b200.sampler import NeighborSampler
    g = make_graph(400, seed=22)
    d = data_all(g)
    n = d.x.shape[0]
    perm = torch.from_numpy(np.random.default_rng(3).permutation(d.edge_index.shape[1]))
    ei = d.edge_index[:, perm].contiguous()
    ref = OracleSampler(ei, [-1] * 2, 50, num_nodes=n)
    smp = NeighborSampler(ei, [-1] * 2, batch_size=50, num_nodes=n, device=DEV)
    for (bs, n_id, adjs), (bs_r, n_id_r, adjs_r) in zip(smp, ref):
        assert torch.equal(n_id.cpu(), n_id_r)
        for (e, e_id, size), (e_r, e_id_r, size_r) in zip(adjs, adjs_r):
            assert torch.equal(e.cpu(), e_r) and torch.equal(e_id.cpu(), e_id_r)
            src = ei[0][e_id_r]
            tgt_l = e_r[1]
            same = tgt_l[1:] == tgt_l[:-1]
            assert bool((src[1:][same] >= src[:-1][same]).all())
