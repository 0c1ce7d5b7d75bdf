"""Generate the golden fixtures that pin the oracle (run in the BUILD container only).

    python tests/golden/make_golden.py

Executes the reference's own, unmodified source from ``/root/reference`` —
``learning/surfaceNetStaticEdgeFilters.py``, ``learning/surfaceNetUpdatedEdgeFilters.py``,
``learning/runModel.py`` — over the third-party stand-ins of ``pyg_shim.py`` and writes small
``.npz`` fixtures next to this file.  ``/root/reference`` does not exist on the GPU box; the
tests only read the committed fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import pyg_shim  # noqa: E402

pyg_shim.install()
sys.path.insert(0, os.path.join(REF, "learning"))
import surfaceNetStaticEdgeFilters as ref_static  # noqa: E402
import surfaceNetUpdatedEdgeFilters as ref_updated  # noqa: E402
import runModel as ref_rm  # noqa: E402

from oracle import graph as og  # noqa: E402
from oracle.static_model import NeighborSampler, make_clf, to_attr  # noqa: E402

torch.set_num_threads(1)
torch.use_deterministic_algorithms(True)


def sd_to_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def small_graph(n_points, seed):
    pts = og.scan_like_points(n_points, seed=seed)
    adj, infinite, cen, tets = og.delaunay_graph(pts)
    N = infinite.shape[0]
    x, ea, y = og.synthetic_features(N, infinite, seed=seed + 1)
    return adj, infinite, cen, x, ea, y


def data_all(adj, x, ea, y):
    return to_attr(dict(x=torch.from_numpy(x), edge_attr=torch.from_numpy(ea), y=torch.from_numpy(y),
                        edge_index=torch.from_numpy(adj.T.astype(np.int64)).contiguous()))


def main():
    out = {}
    # ------------------------------------------------------------------ graph + features
    adj, infinite, cen, x, ea, y = small_graph(220, seed=3)
    N = infinite.shape[0]
    out.update(adj=adj, infinite=infinite, cen=cen, x=x, ea=ea, y=y)
    d = data_all(adj, x, ea, y)

    # ------------------------------------------------------------------ Static, kf96 weights, eval
    clf = make_clf()
    kf96 = torch.load(os.path.join(REF, "data/models/kf96/model_best.ptm"), map_location="cpu")
    m = ref_static.SurfaceNet(clf=clf)
    m.load_state_dict(kf96, strict=True)
    m.eval()
    with torch.no_grad():
        out["kf96_inference_layer"] = m.inference_layer(d).numpy()
        ldr = NeighborSampler(d.edge_index, sizes=[-1] * 4, batch_size=256, num_nodes=N)
        out["kf96_inference_batch_layer"] = m.inference_batch_layer(d, ldr).numpy()
        ldr1 = NeighborSampler(d.edge_index, sizes=[-1], batch_size=256, num_nodes=N)
        out["kf96_inference_layer_batch"] = m.inference_layer_batch(d, ldr1).numpy()
    np.savez_compressed(os.path.join(HERE, "kf96_state.npz"), **sd_to_np(kf96, ""))

    # loss on the eval logits through the reference Trainer (runModel.py:438-445)
    tr = ref_rm.Trainer(m)
    d.batch_x, d.batch_gt, d.batch_adjs = d.x, d.y, []
    for cell_norm in (None, "sqrt", "log"):
        clf.regularization.cell_norm = cell_norm
        metrics = ref_rm.Metrics()
        loss = tr.calcLossAndOA(torch.from_numpy(out["kf96_inference_layer"]), None, d, clf, metrics)
        out["kf96_loss_%s" % cell_norm] = np.float32(loss.item())
        out["kf96_oa_count"] = np.int64(metrics.OA_sum)
    clf.regularization.cell_norm = None
    # regulariser value (runModel.py:109-160, unbatched branch)
    metrics = ref_rm.Metrics()
    d.edge_index_backup = d.edge_index
    reg = tr.calcRegularization(torch.from_numpy(out["kf96_inference_layer"]), d, clf, metrics)
    out["kf96_reg"] = np.float32(reg.item())

    # ------------------------------------------------------------------ Static, small widths, one train step
    for tag, edge_convs, decoder, norm in (("a", 1, 2, "b"), ("b", 0, 1, "b"), ("c", 2, 2, "l")):
        clf_t = make_clf(convs=(16, 32, 32, 32), edge_convs=edge_convs, decoder=decoder, normalization=norm)
        if not edge_convs:
            clf_t.temp.num_edge_features = None
        torch.manual_seed(0)
        mt = ref_static.SurfaceNet(clf=clf_t)
        out.update(sd_to_np(mt.state_dict(), "train_%s_init." % tag))
        # sampled closure: 96 seeds, num_hops + additional hop (run.py:68,72-74)
        seeds = torch.arange(40, 136)
        smp = NeighborSampler(d.edge_index, sizes=[-1] * 5, batch_size=96, node_idx=seeds, num_nodes=N)
        bs, n_id, adjs = next(iter(smp))
        data = to_attr(dict(all=d, batch_n_id=n_id, batch_adjs=adjs))
        out["train_%s_n_id" % tag] = n_id.numpy()
        for li, (ei, e_id, size) in enumerate(adjs):
            out["train_%s_adj%d_ei" % (tag, li)] = ei.numpy()
            out["train_%s_adj%d_eid" % (tag, li)] = e_id.numpy()
            out["train_%s_adj%d_size" % (tag, li)] = np.asarray(size, dtype=np.int64)
        clf_t.training.metrics = ref_rm.Metrics()
        clf_t.model.edge_prediction = 0
        opt = torch.optim.Adam(mt.parameters(), lr=clf_t.training.learning_rate)
        # forward + loss + backward + Adam exactly as Trainer.train (runModel.py:264-282),
        # with the intermediate values captured
        mt.train()
        logits = mt(data)
        n_sup = adjs[mt.num_layers - 1][2][1]
        data.batch_x = d.x[n_id[:n_sup]]
        data.batch_gt = d.y[n_id[:n_sup]]
        loss = ref_rm.Trainer(mt).calcLossAndOA(logits, None, data, clf_t, clf_t.training.metrics)
        opt.zero_grad()
        loss.backward()
        out["train_%s_logits" % tag] = logits.detach().numpy()
        out["train_%s_loss" % tag] = np.float32(loss.item())
        for k, p in mt.named_parameters():
            out["train_%s_grad.%s" % (tag, k)] = p.grad.detach().numpy().copy()
        opt.step()
        out.update(sd_to_np(mt.state_dict(), "train_%s_after." % tag))

    # ------------------------------------------------------------------ Updated edge filters, forward
    # "sage+" adds out_net; its backward raises in the reference itself (F.relu output is
    # modified in place by out_net's ReLU(True), Updated:245-247), so gradients are pinned on
    # the plain "sage" variant and "sage+" is pinned forward-only.
    smp = NeighborSampler(d.edge_index, sizes=[-1] * 4, batch_size=96, node_idx=torch.arange(40, 136), num_nodes=N)
    bs, n_id, adjs = next(iter(smp))
    out["upd_n_id"] = n_id.numpy()
    for tag, name in (("upd", "sage"), ("updp", "sage+")):
        clf_u = to_attr(dict(training=dict(model_params=[16, 32, 32, 32], model_name=name),
                             features=dict(normalization_feature=1, keep_normalization_feature=0),
                             temp=dict(device="cpu")))
        torch.manual_seed(0)
        mu = ref_updated.SurfaceNet(28, clf_u)
        out.update(sd_to_np(mu.state_dict(), tag + "_init."))
        du = to_attr(dict(x=d.x, edge_attr=d.edge_attr, n_id=n_id, adjs=adjs))
        # the reference calls torch.cuda.empty_cache() (a no-op without CUDA)
        yu = mu(du)
        out[tag + "_logits"] = yu.detach().numpy()
        if name == "sage":
            yu.square().sum().backward()
            for k, p in mu.named_parameters():
                out[tag + "_grad.%s" % k] = p.grad.detach().numpy().copy()

    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    print("wrote golden_small.npz with %d arrays, N=%d" % (len(out), N))


if __name__ == "__main__":
    main()
