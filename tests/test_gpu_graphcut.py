"""Graph-cut regularisation on the device (generate_mesh.py:15-58) against the oracle: the labelling returned is a
global minimiser of the reference's energy (same integer energy as the s-t cut / the converged alpha-expansion)."""
import numpy as np
import pytest
import torch

from oracle import graphcut as ogc
from oracle.static_model import to_attr
from tests.test_graphcut_cpu import _instance

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n_points,seed,w", [(60, 0, 1), (60, 1, 25), (400, 2, 120), (400, 3, 400), (3000, 4, 60), (3000, 5, 0)])
def test_device_cut_is_a_global_minimum(n_points, seed, w):
    from dgnn_b200.generate_mesh import CutGraph, cut_energy, min_cut_labels
    z, edges, uw = _instance(n_points, seed)
    cost = ogc.data_costs(z, uw)
    best = ogc.min_cut(cost, edges, w)
    e_data, e_smooth = ogc.energy(best, cost, edges, w)
    g = CutGraph(edges, z.shape[0], DEV)
    # facet table is symmetric with the right reverse slots
    nbr, rs = g.nbr.cpu().numpy(), g.rslot.cpu().numpy()
    c, k = np.nonzero(nbr >= 0)
    assert np.array_equal(nbr[nbr[c, k], rs[c, k]], c) and (nbr >= 0).sum() == 2 * len(edges)
    lab, stats = min_cut_labels(g, torch.from_numpy(z), uw, w, return_stats=True)
    lab_np = lab.cpu().numpy().astype(np.int64)
    assert set(np.unique(lab_np)) <= {0, 1}
    d, s = ogc.energy(lab_np, cost, edges, w)
    assert d + s == e_data + e_smooth, (d + s, e_data + e_smooth, stats)
    assert cut_energy(g, torch.from_numpy(z), lab, uw, w) == (d, s)          # the device's own energy evaluation
    if w == 0:                                                              # no smoothness: the cut is the argmin per cell
        assert np.array_equal(lab_np[cost[:, 0] != cost[:, 1]], np.argmin(cost, axis=1)[cost[:, 0] != cost[:, 1]])


def test_graph_cut_keeps_the_reference_signature():
    from dgnn_b200.generate_mesh import graph_cut
    z, edges, uw = _instance(500, 7)
    clf = to_attr(dict(graph_cut=dict(unary_weight=uw, binary_weight=80, binary_type=None), temp=dict(device=DEV)))
    labels0 = np.argmax(z, axis=1)
    pred = torch.from_numpy(z.copy())
    out = graph_cut(labels0, pred, edges, clf)
    assert out.dtype == np.int64 and out.shape == labels0.shape
    assert np.array_equal(pred.numpy(), z)                        # not modified in place
    cost = ogc.data_costs(z, uw)
    assert sum(ogc.energy(out, cost, edges, 80)) == sum(ogc.energy(ogc.alpha_expansion(labels0, cost, edges, 80), cost, edges, 80))
    # stronger smoothing never increases the number of interface facets
    cuts = []
    for bw in (0, 40, 400):
        clf.graph_cut.binary_weight = bw
        lab = graph_cut(labels0, pred, edges, clf)
        cuts.append(int((lab[edges[:, 0]] != lab[edges[:, 1]]).sum()))
    assert cuts[0] >= cuts[1] >= cuts[2]


def test_labels_cut_and_interface_chain():
    """generate_mesh.py:75-105 on the device: finite-cell labels -> graph cut -> interface facets."""
    from dgnn_b200.generate_mesh import cell_labels, graph_cut, interface_facet_ids
    from oracle import graph as og
    rng = np.random.default_rng(11)
    adj, infinite, cen, tets = og.delaunay_graph(og.random_points(300, seed=11))
    n = infinite.shape[0]
    z = torch.from_numpy((rng.standard_normal((n, 2)) * 2).astype(np.float32)).to(DEV)
    inf_t = torch.from_numpy(infinite.astype(bool))
    lab0 = cell_labels(z, inf_t)
    fin = np.nonzero(infinite == 0)[0]
    assert np.array_equal(lab0.cpu().numpy(), og.labels_from_logits(z.cpu().numpy()[fin]))
    # nfacets in finite numbering, -1 = infinite
    remap = -np.ones(n, dtype=np.int64); remap[fin] = np.arange(len(fin))
    a, b = adj[:, 0].astype(np.int64), adj[:, 1].astype(np.int64)
    once = (a < b) & ~((infinite[a] == 1) & (infinite[b] == 1))
    nfacets = np.stack([remap[a[once]], remap[b[once]]], axis=1)
    gc_edges = nfacets[(nfacets >= 0).all(axis=1)]
    clf = to_attr(dict(graph_cut=dict(unary_weight=100.0, binary_weight=50, binary_type=None), temp=dict(device=DEV)))
    lab = graph_cut(lab0.cpu().numpy(), z[~inf_t.to(DEV)].cpu(), gc_edges, clf)
    ids = interface_facet_ids(torch.from_numpy(lab.astype(np.uint8)).to(DEV), nfacets).cpu().numpy()
    assert np.array_equal(ids, og.interface_facets(lab, nfacets))
