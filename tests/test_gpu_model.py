"""GPU parity of the Static edge-filter network: CUDA path (through the C ABI) vs the CPU oracle
and the golden vectors.  Tolerance (north_star): fp32 logits within 1e-4 relative to the logit
scale; labels identical off exact ties."""
import numpy as np
import pytest
import torch

from oracle import trainer as otr
from oracle.static_model import NeighborSampler, SurfaceNet as OracleNet, make_clf, to_attr
from tests.helpers import data_all, full_batch, grad_close, labels_equal_off_ties, logits_close, make_graph

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cuda_net(clf_kwargs, state):
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    net = SurfaceNet(make_clf(device=DEV, **clf_kwargs))
    net.load_state_dict(state, strict=True)
    return net.to(DEV)


def golden_data(golden):
    return to_attr(dict(x=torch.from_numpy(golden["x"]), edge_attr=torch.from_numpy(golden["ea"]),
                        y=torch.from_numpy(golden["y"]),
                        edge_index=torch.from_numpy(golden["adj"].T.astype(np.int64)).contiguous()))


def test_kf96_inference_matches_reference_golden(golden, kf96_state):
    net = cuda_net({}, kf96_state).eval()
    d = golden_data(golden)
    z = net.inference_layer(d)
    assert z.device.type == "cuda" and z.shape == (d.x.shape[0], 2)
    err, ok = logits_close(z.cpu().numpy(), golden["kf96_inference_layer"])
    assert ok, err
    flips, ties = labels_equal_off_ties(z.cpu().numpy(), golden["kf96_inference_layer"])
    assert flips == 0
    # the two batched schedules of the reference return the same logits
    for name in ("inference_layer_batch", "inference_batch_layer"):
        zb = getattr(net, name)(d, None)
        err, ok = logits_close(zb.cpu().numpy(), golden["kf96_" + name])
        assert ok, (name, err)


@pytest.mark.parametrize("order", ["pos", "rcm"])
def test_inference_medium_graph_vs_oracle(kf96_state, order):
    g = make_graph(6000, seed=11)
    d = data_all(g, with_pos=(order == "pos"))
    ref = OracleNet(make_clf()); ref.load_state_dict(kf96_state); ref.eval()
    with torch.no_grad():
        zr = ref.inference_layer(d).numpy()
    net = cuda_net({}, kf96_state).eval()
    z = net.inference_layer(d).cpu().numpy()
    err, ok = logits_close(z, zr)
    assert ok, err
    flips, ties = labels_equal_off_ties(z, zr)
    assert flips == 0


def _train_compare(clf_kwargs, data, d, n_sup_rows, seed=0, loss_kwargs=None):
    from dgnn_b200 import runModel as rm
    torch.manual_seed(seed)
    ref = OracleNet(make_clf(**clf_kwargs))
    # make the norm affine non-trivial so its gradients are exercised
    with torch.no_grad():
        for k, p in ref.named_parameters():
            if "norm" in k or k.startswith("decoder.1"):
                p.add_(0.3 * torch.randn_like(p))
    net = cuda_net(clf_kwargs, ref.state_dict())
    ref.train(); net.train()
    clf = make_clf(device=DEV, **clf_kwargs)
    if loss_kwargs:
        clf.regularization.update(loss_kwargs)
    zr = ref(data)
    gt = d.y[data.batch_n_id[:n_sup_rows]]
    bx = d.x[data.batch_n_id[:n_sup_rows]]
    lr, _, _ = otr.cell_loss(zr, gt, bx[:, 0], clf.training.loss, clf.regularization.cell_norm,
                             clf.regularization.cell_type)
    lr.backward()
    z = net(data)
    assert z.requires_grad and z.shape == zr.shape
    loss = rm.cell_loss(z, gt, bx, clf)
    loss.backward()
    err, ok = logits_close(z.detach().cpu().numpy(), zr.detach().numpy())
    assert ok, err
    assert abs(loss.item() - lr.item()) <= 2e-5 * max(1.0, abs(lr.item())), (loss.item(), lr.item())
    refp = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        e, tol = grad_close(p.grad, refp[k].grad)
        assert e <= tol, (k, e, tol)
    # running statistics were updated like BatchNorm1d does
    for k, b in net.named_buffers():
        rb = dict(ref.named_buffers())[k]
        np.testing.assert_allclose(b.cpu().numpy(), rb.numpy(), rtol=2e-4, atol=1e-6, err_msg=k)
    return net, ref


@pytest.mark.parametrize("loss", ["bce", "mse"])
def test_train_single_logit_losses(loss):
    """Row S: the bce / mse branches of calcLossAndOA (one logit per cell, runModel.py:181-188) through a whole
    train step: logits, loss value and every gradient against the oracle."""
    g = make_graph(500, seed=27)
    d = data_all(g)
    # batch_gt columns: inside %, outside %, (unused), graph-cut label
    d.y = torch.cat([d.y, torch.zeros(d.y.shape[0], 1), (d.y[:, :1] > d.y[:, 1:2]).float()], dim=1)
    _train_compare(dict(convs=(16, 32, 32, 32), loss=loss), full_batch(d), d, d.x.shape[0])


def test_train_full_graph_kf96_widths():
    g = make_graph(1200, seed=21)
    d = data_all(g)
    data = full_batch(d)
    _train_compare({}, data, d, d.x.shape[0])


@pytest.mark.parametrize("cell_norm", ["sqrt", "log"])
def test_train_loss_weight_variants(cell_norm):
    g = make_graph(300, seed=22)
    d = data_all(g)
    _train_compare(dict(convs=(16, 32, 32, 32)), full_batch(d), d, d.x.shape[0], loss_kwargs=dict(cell_norm=cell_norm))


@pytest.mark.parametrize("kw", [dict(convs=(16, 32, 32, 32)),
                                dict(convs=(16, 32, 32, 32), edge_convs=0, decoder=1, n_edge_feat=None),
                                dict(convs=(32, 64), decoder=0),
                                dict(convs=(16, 32, 32), normalization="l")])
def test_train_sampled_closure_matches_oracle(golden, kw):
    """Seed-batched training semantics of the reference (run.py:72-74): per-layer bipartite
    sub-graphs with n_src != n_tgt."""
    d = golden_data(golden)
    L = len(kw["convs"])
    smp = NeighborSampler(d.edge_index, [-1] * (L + 1), 96, node_idx=torch.arange(40, 136), num_nodes=d.x.shape[0])
    bs, n_id, adjs = next(iter(smp))
    data = to_attr(dict(all=d, batch_n_id=n_id, batch_adjs=adjs))
    n_sup = adjs[L - 1][2][1]
    if kw.get("decoder", 2) == 0:
        pytest.skip("decoder=0 emits hidden features, no kl loss")
    _train_compare(kw, data, d, n_sup)


@pytest.mark.parametrize("tag,kw", [("a", dict(convs=(16, 32, 32, 32))),
                                    ("c", dict(convs=(16, 32, 32, 32), edge_convs=2, decoder=2, normalization="l"))])
def test_golden_train_step(golden, tag, kw):
    """The reference's own train step on a sampled closure (golden cases 'a': shipped options, 'c': two-layer edge
    MLP + graph LayerNorm): logits, loss and every gradient."""
    from dgnn_b200 import runModel as rm
    d = golden_data(golden)
    pre = "train_%s_" % tag
    state = {k[len(pre + "init."):]: torch.from_numpy(v) for k, v in golden.items() if k.startswith(pre + "init.")}
    net = cuda_net(kw, state).train()
    n_id = torch.from_numpy(golden[pre + "n_id"])
    adjs = [(torch.from_numpy(golden[pre + "adj%d_ei" % i]), torch.from_numpy(golden[pre + "adj%d_eid" % i]),
             tuple(int(v) for v in golden[pre + "adj%d_size" % i])) for i in range(5)]
    data = to_attr(dict(all=d, batch_n_id=n_id, batch_adjs=adjs))
    z = net(data)
    n_sup = adjs[3][2][1]
    clf = make_clf(device=DEV, **kw)
    loss = rm.cell_loss(z, d.y[n_id[:n_sup]], d.x[n_id[:n_sup]], clf)
    loss.backward()
    err, ok = logits_close(z.detach().cpu().numpy(), golden[pre + "logits"])
    assert ok, err
    assert abs(loss.item() - float(golden[pre + "loss"])) <= 2e-5
    for k, p in net.named_parameters():
        e, tol = grad_close(p.grad, torch.from_numpy(golden[pre + "grad." + k]))
        assert e <= tol, (k, e, tol)


def test_adam_matches_torch():
    from dgnn_b200.runModel import Adam
    torch.manual_seed(0)
    ps = [torch.randn(7, 5), torch.randn(33), torch.randn(128, 64)]
    a = [p.clone().cuda().requires_grad_() for p in ps]
    b = [p.clone().requires_grad_() for p in ps]
    oa, ob = Adam(a, lr=0.005), torch.optim.Adam(b, lr=0.005)
    for step in range(5):
        for x, y in zip(a, b):
            gr = torch.randn_like(y)
            y.grad = gr.clone(); x.grad = gr.cuda()
        oa.step(); ob.step()
    for x, y in zip(a, b):
        np.testing.assert_allclose(x.detach().cpu().numpy(), y.detach().numpy(), rtol=1e-5, atol=1e-7)


def test_training_reduces_loss_and_matches_oracle_trajectory():
    """Five Adam steps on a small graph: the loss trajectory follows the oracle's."""
    from dgnn_b200 import runModel as rm
    g = make_graph(300, seed=31)
    d = data_all(g)
    data = full_batch(d)
    kw = dict(convs=(16, 32, 32, 32))
    torch.manual_seed(0)
    ref = OracleNet(make_clf(**kw))
    net = cuda_net(kw, ref.state_dict())
    clf = make_clf(device=DEV, **kw)
    oa, ob = rm.Adam(net.parameters(), lr=0.005), torch.optim.Adam(ref.parameters(), lr=0.005)
    la, lb = [], []
    for _ in range(5):
        net.train(); ref.train()
        loss = rm.cell_loss(net(data), d.y, d.x, clf)
        oa.zero_grad(); loss.backward(); oa.step()
        lr, _, _ = otr.cell_loss(ref(data), d.y, d.x[:, 0])
        ob.zero_grad(); lr.backward(); ob.step()
        la.append(loss.item()); lb.append(lr.item())
    assert la[-1] < la[0]
    np.testing.assert_allclose(la, lb, rtol=2e-3)


def test_unsupported_options_fail_loudly():
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    g = make_graph(60, seed=1)
    net = SurfaceNet(make_clf(device=DEV, convs=(18, 32, 32, 32))).to(DEV).eval()   # hidden width not a multiple of 4
    with pytest.raises(NotImplementedError):
        net.inference_layer(data_all(g))


@pytest.mark.parametrize("norm", ["b", "l"])
def test_train_two_layer_edge_mlp(norm):
    """Row C with ``edge_convs == 2`` (Linear -> norm over the edges -> ReLU -> Linear as the edge filter): a whole
    train step (logits, loss, every gradient, running statistics) against the oracle, and eval-mode inference."""
    g = make_graph(500, seed=31)
    d = data_all(g)
    kw = dict(convs=(16, 32, 32, 32), edge_convs=2, normalization=norm)
    net, ref = _train_compare(kw, full_batch(d), d, d.x.shape[0])
    if norm == "b":
        net.eval(); ref.eval()
        with torch.no_grad():
            z, zr = net.inference_layer(d), ref.inference_layer(d)
        err, ok = logits_close(z.cpu().numpy(), zr.numpy())
        assert ok, err


def test_updated_edge_filters_forward_matches_reference_golden(golden):
    """Row U: the Updated-edge-filter variant (forward) against the reference's own output."""
    from dgnn_b200.surfaceNetUpdatedEdgeFilters import SurfaceNet as UpdNet
    d = golden_data(golden)
    n_id = torch.from_numpy(golden["upd_n_id"])
    bs, n_id2, adjs = next(iter(NeighborSampler(d.edge_index, [-1] * 4, 96, node_idx=torch.arange(40, 136),
                                                num_nodes=d.x.shape[0])))
    assert torch.equal(n_id, n_id2)
    for tag, name in (("upd", "sage"), ("updp", "sage+")):
        clf = to_attr(dict(training=dict(model_params=[16, 32, 32, 32], model_name=name),
                           features=dict(normalization_feature=1, keep_normalization_feature=0),
                           temp=dict(device=DEV)))
        m = UpdNet(28, clf)
        m.load_state_dict({k[len(tag) + 6:]: torch.from_numpy(v) for k, v in golden.items() if k.startswith(tag + "_init.")},
                          strict=True)
        m.to(DEV)
        y = m(to_attr(dict(x=d.x, edge_attr=d.edge_attr, n_id=n_id, adjs=adjs)))
        err, ok = logits_close(y.detach().cpu().numpy(), golden[tag + "_logits"])
        assert ok, (name, err)


def _upd_clf(name, device, params=(16, 32, 32, 32)):
    return to_attr(dict(training=dict(model_params=list(params), model_name=name),
                        features=dict(normalization_feature=1, keep_normalization_feature=0),
                        temp=dict(device=device)))


def test_updated_edge_filters_gradients_match_reference_golden(golden):
    """Row U, training: gradients of every parameter of the Updated-edge-filter model ("sage") against the
    reference's own autograd on the sampled 4-hop closure of the golden graph (loss = sum(out^2))."""
    from dgnn_b200.surfaceNetUpdatedEdgeFilters import SurfaceNet as UpdNet
    d = golden_data(golden)
    n_id = torch.from_numpy(golden["upd_n_id"])
    _, _, adjs = next(iter(NeighborSampler(d.edge_index, [-1] * 4, 96, node_idx=torch.arange(40, 136),
                                           num_nodes=d.x.shape[0])))
    m = UpdNet(28, _upd_clf("sage", DEV))
    m.load_state_dict({k[len("upd_init."):]: torch.from_numpy(v) for k, v in golden.items() if k.startswith("upd_init.")},
                      strict=True)
    m.to(DEV)
    y = m(to_attr(dict(x=d.x, edge_attr=d.edge_attr, n_id=n_id, adjs=adjs)))
    assert y.requires_grad
    y.square().sum().backward()
    for k, p in m.named_parameters():
        err, tol = grad_close(p.grad, torch.from_numpy(golden["upd_grad." + k]))
        assert err <= tol, (k, err)


@pytest.mark.parametrize("name", ["sage", "sage+"])
def test_updated_edge_filters_full_graph_training_matches_oracle(name):
    """Row U at cfg4 shape (whole graph as the batch, kf96-like widths): logits and every gradient against the
    oracle; the "+" head, whose backward the reference itself cannot run, is pinned here."""
    from dgnn_b200.surfaceNetUpdatedEdgeFilters import SurfaceNet as UpdNet
    from oracle.updated_model import SurfaceNet as OracleUpd
    gr = make_graph(700, seed=5)
    d = data_all(gr)
    n = d.x.shape[0]
    e = d.edge_index.shape[1]
    adjs = [(d.edge_index, torch.arange(e), (n, n))] * 4
    batch = dict(x=d.x, edge_attr=d.edge_attr, n_id=torch.arange(n), adjs=adjs)
    torch.manual_seed(3)
    ref = OracleUpd(28, _upd_clf(name, "cpu", (64, 128, 128, 128)))
    if name == "sage+":     # the in-place ReLUs of the head (Updated:210) break torch's own backward; same math
        ref.out_net[0], ref.out_net[2] = torch.nn.ReLU(False), torch.nn.ReLU(False)
    m = UpdNet(28, _upd_clf(name, DEV, (64, 128, 128, 128)))
    m.load_state_dict(ref.state_dict(), strict=True)
    m.to(DEV)
    w = torch.linspace(0.5, 1.5, n)[:, None]
    y_ref = ref(to_attr(batch))
    (y_ref.square() * w).sum().backward()
    y = m(to_attr(batch))
    err, ok = logits_close(y.detach().cpu().numpy(), y_ref.detach().numpy())
    assert ok, err
    (y.square() * w.to(DEV)).sum().backward()
    gref = dict(ref.named_parameters())
    for k, p in m.named_parameters():
        err, tol = grad_close(p.grad, gref[k].grad)
        assert err <= tol, (name, k, err)
    with torch.no_grad():
        y2 = m(to_attr(batch))
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("n_points", [5, 6, 9])
def test_tiny_graphs(n_points):
    """Smallest tetrahedralisations (one to a handful of tetrahedra plus their infinite cells; far fewer rows than one
    128-cell tile): inference and a train step still match the oracle."""
    g = make_graph(n_points, seed=40 + n_points, scan_like=False)
    d = data_all(g)
    kw = dict(convs=(16, 32, 32, 32))
    net, ref = _train_compare(kw, full_batch(d), d, d.x.shape[0])
    net.eval(); ref.eval()
    with torch.no_grad():
        z, zr = net.inference_layer(d), ref.inference_layer(d)
    err, ok = logits_close(z.cpu().numpy(), zr.numpy())
    assert ok, err


def _with_self_loops(d):
    """``add_self_loops(edge_index)[0]`` (run.py:70-71,215-216): one edge i -> i per node behind the facet edges."""
    n = d.x.shape[0]
    own = torch.arange(n, dtype=torch.int64)
    # edge_attr keeps its 4 N rows: the reference samples with return_e_id = model.edge_convs = 0 and never reads it
    return to_attr(dict(x=d.x, y=d.y, edge_attr=d.edge_attr,
                        edge_index=torch.cat([d.edge_index, torch.stack([own, own])], dim=1)))


def _full_batch_no_eid(d, n_layers_plus=5):
    n = d.x.shape[0]
    return to_attr(dict(all=d, batch_n_id=torch.arange(n), batch_adjs=[(d.edge_index, None, (n, n))] * n_layers_plus))


SL_KW = dict(convs=(16, 32, 32, 32), edge_convs=0, decoder=1, n_edge_feat=None)


def test_self_loops_whole_graph_train_step_and_inference():
    """``graph.self_loops: 1`` with ``model.edge_convs: 0`` (run.py:70-71,215-216): every cell is its own fifth
    in-neighbour.  Whole-graph train step (logits, loss, every gradient) and inference against the oracle."""
    g = make_graph(700, seed=33)
    d = _with_self_loops(data_all(g))
    net, ref = _train_compare(SL_KW, _full_batch_no_eid(d), d, d.x.shape[0])
    net.eval(); ref.eval()
    with torch.no_grad():
        z, zr = net.inference_layer(d).cpu().numpy(), ref.inference_layer(d).numpy()
    err, ok = logits_close(z, zr)
    assert ok, err


def test_self_loops_sampled_closure_and_device_sampler():
    """Seed-batch training on a self-looped graph: the closure sampled by the oracle's restatement of PyG's sampler
    (five in-edges per target), and the device sampler yielding the same nodes and the same edges per hop."""
    from dgnn_b200.sampler import NeighborSampler as DeviceSampler
    g = make_graph(500, seed=34)
    d = _with_self_loops(data_all(g))
    L = len(SL_KW["convs"])
    seeds = torch.arange(30, 126)
    smp = NeighborSampler(d.edge_index, [-1] * (L + 1), 96, node_idx=seeds, num_nodes=d.x.shape[0])
    bs, n_id, adjs = next(iter(smp))
    adjs = [(a[0], None, a[2]) for a in adjs]              # return_e_id = model.edge_convs = 0 (run.py:74)
    data = to_attr(dict(all=d, batch_n_id=n_id, batch_adjs=adjs))
    _train_compare(SL_KW, data, d, adjs[L - 1][2][1])
    dsm = DeviceSampler(d.edge_index, [-1] * (L + 1), node_idx=seeds, num_nodes=d.x.shape[0], batch_size=96,
                        return_e_id=False, device=DEV)
    assert dsm.self_loops
    bs2, n_id2, adjs2 = next(iter(dsm))
    assert bs2 == bs and torch.equal(n_id2.cpu(), n_id)
    for a, b in zip(adjs, adjs2):
        assert tuple(a[2]) == tuple(b.size)
        ka = sorted(map(tuple, a[0].t().tolist())); kb = sorted(map(tuple, b.edge_index.t().cpu().tolist()))
        assert ka == kb
