"""The oracle against the golden vectors produced by the reference's own source
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import trainer as otr
from oracle.static_model import NeighborSampler, SurfaceNet, make_clf, to_attr
from oracle.updated_model import SurfaceNet as UpdatedNet

TOL = dict(rtol=1e-5, atol=2e-6)


def sub_state(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def make_data(g):
    return to_attr(dict(x=torch.from_numpy(g["x"]), edge_attr=torch.from_numpy(g["ea"]),
                        y=torch.from_numpy(g["y"]),
                        edge_index=torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()))


def test_kf96_loads_strict_and_matches_reference_inference(golden, kf96_state):
    m = SurfaceNet(make_clf())
    m.load_state_dict(kf96_state, strict=True)
    assert sum(v.numel() for v in m.state_dict().values()) == 103699
    m.eval()
    d = make_data(golden)
    N = d.x.shape[0]
    with torch.no_grad():
        z = m.inference_layer(d)
        np.testing.assert_allclose(z.numpy(), golden["kf96_inference_layer"], **TOL)
        zb = m.inference_batch_layer(d, NeighborSampler(d.edge_index, [-1] * 4, 256, num_nodes=N))
        np.testing.assert_allclose(zb.numpy(), golden["kf96_inference_batch_layer"], **TOL)
        zl = m.inference_layer_batch(d, NeighborSampler(d.edge_index, [-1], 256, num_nodes=N))
        np.testing.assert_allclose(zl.numpy(), golden["kf96_inference_layer_batch"], **TOL)
    # the three reference schedules agree with each other
    np.testing.assert_allclose(zb.numpy(), z.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(zl.numpy(), z.numpy(), rtol=1e-4, atol=1e-5)


def test_loss_and_regulariser_match_reference(golden):
    d = make_data(golden)
    z = torch.from_numpy(golden["kf96_inference_layer"])
    for cn in (None, "sqrt", "log"):
        loss, _, _ = otr.cell_loss(z, d.y, d.x[:, 0], "kl", cn, "vol")
        np.testing.assert_allclose(loss.item(), golden["kf96_loss_%s" % cn], rtol=1e-6)
    assert otr.overall_accuracy_count(z, d.y) == int(golden["kf96_oa_count"])
    reg = otr.edge_regularization(z, d.edge_index, 0.4)
    np.testing.assert_allclose(reg.item(), golden["kf96_reg"], rtol=1e-6)


def _train_case(golden, tag, edge_convs, decoder, norm):
    clf = make_clf(convs=(16, 32, 32, 32), edge_convs=edge_convs, decoder=decoder, normalization=norm)
    if not edge_convs:
        clf.temp.num_edge_features = None
    m = SurfaceNet(clf)
    m.load_state_dict(sub_state(golden, "train_%s_init." % tag), strict=True)
    d = make_data(golden)
    n_id = torch.from_numpy(golden["train_%s_n_id" % tag])
    adjs = []
    for li in range(5):
        adjs.append((torch.from_numpy(golden["train_%s_adj%d_ei" % (tag, li)]),
                     torch.from_numpy(golden["train_%s_adj%d_eid" % (tag, li)]),
                     tuple(int(v) for v in golden["train_%s_adj%d_size" % (tag, li)])))
    # the oracle sampler reproduces the fixture's closure
    bs, n_id2, adjs2 = next(iter(NeighborSampler(d.edge_index, [-1] * 5, 96, node_idx=torch.arange(40, 136),
                                                 num_nodes=d.x.shape[0])))
    assert torch.equal(n_id, n_id2)
    for a, b in zip(adjs, adjs2):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[2] == b[2]
    data = to_attr(dict(all=d, batch_n_id=n_id, batch_adjs=adjs))
    opt = torch.optim.Adam(m.parameters(), lr=clf.training.learning_rate)
    m.train()
    logits = m(data)
    n_sup = adjs[m.num_layers - 1][2][1]
    loss, _, _ = otr.cell_loss(logits, d.y[n_id[:n_sup]], d.x[n_id[:n_sup], 0], "kl", None, "vol")
    opt.zero_grad()
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), golden["train_%s_logits" % tag], **TOL)
    np.testing.assert_allclose(loss.item(), golden["train_%s_loss" % tag], rtol=1e-6)
    for k, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), golden["train_%s_grad.%s" % (tag, k)], rtol=1e-4, atol=1e-7,
                                   err_msg=k)
    opt.step()
    grads = {k: p.grad for k, p in m.named_parameters()}
    for k, v in m.state_dict().items():
        # a bias feeding a BatchNorm has an analytically zero gradient; Adam turns its
        # rounding noise into +-lr steps, so those entries are not comparable.
        if k in grads and grads[k].abs().max() < 1e-6:
            continue
        np.testing.assert_allclose(v.numpy(), golden["train_%s_after.%s" % (tag, k)], rtol=1e-4, atol=2e-5,
                                   err_msg=k)


def test_train_step_edge_convs1_decoder2_bn(golden):
    _train_case(golden, "a", 1, 2, "b")


def test_train_step_edge_convs0_decoder1_bn(golden):
    _train_case(golden, "b", 0, 1, "b")


def test_train_step_edge_convs2_decoder2_layernorm(golden):
    _train_case(golden, "c", 2, 2, "l")


def test_updated_edge_filters_forward_and_grads(golden):
    d = make_data(golden)
    n_id = torch.from_numpy(golden["upd_n_id"])
    bs, n_id2, adjs = next(iter(NeighborSampler(d.edge_index, [-1] * 4, 96, node_idx=torch.arange(40, 136),
                                                num_nodes=d.x.shape[0])))
    assert torch.equal(n_id, n_id2)
    for tag, name in (("upd", "sage"), ("updp", "sage+")):
        clf = to_attr(dict(training=dict(model_params=[16, 32, 32, 32], model_name=name),
                           features=dict(normalization_feature=1, keep_normalization_feature=0),
                           temp=dict(device="cpu")))
        m = UpdatedNet(28, clf)
        m.load_state_dict(sub_state(golden, tag + "_init."), strict=True)
        y = m(to_attr(dict(x=d.x, edge_attr=d.edge_attr, n_id=n_id, adjs=adjs)))
        np.testing.assert_allclose(y.detach().numpy(), golden[tag + "_logits"], **TOL)
        if name == "sage":
            y.square().sum().backward()
            for k, p in m.named_parameters():
                np.testing.assert_allclose(p.grad.numpy(), golden["upd_grad.%s" % k], rtol=1e-4, atol=1e-6,
                                           err_msg=k)
