"""The loader oracle (per-graph StandardScaler) against the installed scikit-learn, and the adjacency -> edge_index
convention of ``readAdjacencies_bin``."""
import numpy as np
import torch
from sklearn.preprocessing import StandardScaler

from oracle import graph as og
from oracle.loader import standardize


def _features(n=4000, c=29, seed=0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, c)) * rng.uniform(0.01, 50, c) + rng.uniform(-20, 20, c)).astype(np.float32)
    x[:, 5] = 3.25                       # constant column: left unscaled
    x[:, 7] = 0.0
    x[rng.random(n) < 0.9, 9] = 1.0      # "feat-like" sparsity (most entries at the column mode)
    return x


def test_oracle_standardize_matches_sklearn():
    x = _features()
    ref = StandardScaler().fit_transform(x.astype(np.float64)).astype(np.float32)
    np.testing.assert_allclose(standardize(x), ref, rtol=2e-6, atol=2e-6)
    ref1 = x.copy()
    ref1[:, 1:] = StandardScaler().fit_transform(x[:, 1:].astype(np.float64)).astype(np.float32)
    np.testing.assert_allclose(standardize(x, skip_first=True), ref1, rtol=2e-6, atol=2e-6)
    assert np.array_equal(standardize(x, skip_first=True)[:, 0], x[:, 0])


def test_edge_index_from_adjacencies_convention():
    from dgnn_b200.data import edge_index_from_adjacencies
    adj, infinite, cen, _ = og.delaunay_graph(og.random_points(40, seed=2))
    ei = edge_index_from_adjacencies(adj)
    n = infinite.shape[0]
    assert ei.dtype == torch.int64 and tuple(ei.shape) == (2, 4 * n)
    assert torch.equal(ei[0], torch.arange(n).repeat_interleave(4))          # row 4i+k is owned by cell i
    assert np.array_equal(ei[1].numpy(), adj[:, 1])
