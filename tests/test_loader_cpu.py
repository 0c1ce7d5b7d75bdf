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


def _loader_cases():
    import importlib.util, os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_loader_golden", os.path.join(here, "make_loader_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m, dict(np.load(os.path.join(here, "loader_golden.npz")))


def test_oracle_loader_matches_the_reference_loader_golden():
    """oracle.loader.load_graph against the outputs of the reference's unmodified processing/data.py (column order,
    dropped statistics, regularisation column, the three scalers)."""
    import os
    from oracle.loader import load_graph
    m, gold = _loader_cases()
    base = os.path.join(m.SCENE, "7")
    for tag, c in m.CONFIGS.items():
        r = load_graph(base, m.make_clf(c))
        assert r["node_names"] == list(gold[tag + "_node_names"]), tag
        np.testing.assert_allclose(r["features"], gold[tag + "_features"], rtol=3e-7, atol=1e-9, err_msg=tag)
        assert np.array_equal(r["edge_lists"], gold[tag + "_edge_lists"])
        assert np.array_equal(r["gt"], gold[tag + "_gt"]) and np.array_equal(r["infinite"], gold[tag + "_infinite"])
        assert abs(r["mean_edge"] - float(gold[tag + "_mean_edge"])) < 1e-15
        if c["edge_convs"]:
            assert r["edge_names"] == list(gold[tag + "_edge_names"]), tag
            np.testing.assert_allclose(r["edge_features"], gold[tag + "_edge_features"], rtol=3e-7, atol=1e-9, err_msg=tag)


def test_loader_reproduces_the_reference_errors_on_cpu_side_logic():
    """Dropping a statistic from the facet / edge features calls .drop on an NpzFile in the reference (AttributeError)."""
    import os
    import pytest
    from dgnn_b200.data import _assemble, dataLoader
    m, _ = _loader_cases()
    base = os.path.join(m.SCENE, "7")
    with pytest.raises(AttributeError):
        _assemble(base, ["shape", "facet", "min", "max", "sum"], dataLoader.NODE_SPEC, "node")
    with pytest.raises(AttributeError):
        _assemble(base, ["shape", "vertex", "count", "min", "max"], dataLoader.EDGE_SPEC, "edge")
    names, cols, extra = _assemble(base, ["shape", "vertex", "facet", "count", "min", "max", "sum"], dataLoader.NODE_SPEC, "node")
    assert len(names) == 28 and names[:4] == ["radius", "vol", "longest_edge", "shortest_edge"] and "mean_edge" in extra
