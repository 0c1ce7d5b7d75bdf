"""The graph-cut oracle: the restated alpha-expansion of gco ends at the global minimum that a single
s-t cut gives, from any initial labelling - the property the device implementation relies on."""
import numpy as np
import pytest

from oracle import graph as og
from oracle import graphcut as gc


def _instance(n_points, seed, uw=100.0):
    rng = np.random.default_rng(seed)
    adj, infinite, cen, tets = og.delaunay_graph(og.random_points(n_points, seed=seed))
    fin = np.nonzero(infinite == 0)[0]
    remap = -np.ones(infinite.shape[0], dtype=np.int64); remap[fin] = np.arange(fin.shape[0])
    a, b = adj[:, 0].astype(np.int64), adj[:, 1].astype(np.int64)
    keep = (a < b) & (infinite[a] == 0) & (infinite[b] == 0)            # every finite-finite facet once
    edges = np.stack([remap[a[keep]], remap[b[keep]]], axis=1)
    z = (rng.standard_normal((fin.shape[0], 2)) * 1.5).astype(np.float32)
    return z, edges, uw


@pytest.mark.parametrize("seed,w", [(0, 1), (1, 25), (2, 120), (3, 400)])
def test_alpha_expansion_reaches_the_min_cut_energy(seed, w):
    z, edges, uw = _instance(60, seed)
    cost = gc.data_costs(z, uw)
    best = gc.min_cut(cost, edges, w)
    e_best = sum(gc.energy(best, cost, edges, w))
    rng = np.random.default_rng(seed + 100)
    for init in (np.argmax(-cost, axis=1), np.zeros(len(cost), dtype=np.int64), rng.integers(0, 2, len(cost))):
        lab = gc.alpha_expansion(init, cost, edges, w)
        assert sum(gc.energy(lab, cost, edges, w)) == e_best
    # no labelling found by local perturbation beats it
    for _ in range(50):
        cand = best.copy(); cand[rng.integers(0, len(cand), 3)] ^= 1
        assert sum(gc.energy(cand, cost, edges, w)) >= e_best


def test_data_costs_follow_the_reference_rounding():
    z = np.array([[0.125, -0.375], [2.5, 1.5], [-0.004999, 0.015]], dtype=np.float32)
    c = gc.data_costs(z, 100.0)
    assert c.tolist() == [[-38, 12], [150, 250], [2, 0]]          # columns swapped, round-half-even of the float32 product
