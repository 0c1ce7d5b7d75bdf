"""Loader standardisation on the device against the oracle (itself pinned to scikit-learn on the CPU)."""
import numpy as np
import pytest
import torch

from oracle.loader import standardize

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,c,skip_first", [(5000, 29, True), (20000, 20, False), (7, 4, False), (3000, 130, True)])
def test_standardize_matches_oracle(n, c, skip_first):
    from dgnn_b200.data import standardize_
    rng = np.random.default_rng(n + c)
    x = (rng.standard_normal((n, c)) * rng.uniform(0.01, 50, c) + rng.uniform(-20, 20, c)).astype(np.float32)
    if c > 6:
        x[:, 5] = 3.25
        x[:, 6] = 0.0
    ref = standardize(x, skip_first=skip_first)
    out = standardize_(torch.from_numpy(x).to(DEV), skip_first=skip_first).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=2e-6)
    if skip_first:
        assert np.array_equal(out[:, 0], x[:, 0])
    # run-to-run reproducible (fixed-order reductions)
    out2 = standardize_(torch.from_numpy(x).to(DEV), skip_first=skip_first).cpu().numpy()
    assert np.array_equal(out, out2)


def _cases():
    import importlib.util, os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_loader_golden", os.path.join(here, "make_loader_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m, dict(np.load(os.path.join(here, "loader_golden.npz")))


def _ulp_close(a, b, name):
    """float32 results of a float64 pipeline: identical up to the last bit (different float64 summation order)."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, name
    np.testing.assert_allclose(a, b, rtol=2.4e-7, atol=1e-9, err_msg=name)
    assert (a != b).mean() < 0.02, (name, (a != b).mean())


@pytest.mark.parametrize("tag", ["kf96", "minmax_last", "sum_nocount", "robust_noedge"])
def test_loader_front_end_matches_reference_golden(tag, tmp_path):
    """dgnn_b200.data.dataLoader (device float64 scaling, no pandas / sklearn) against the reference's own loader
    outputs; then the binary cache: second run is a hit and returns the very same tensors."""
    import shutil
    from dgnn_b200.data import dataLoader
    m, gold = _cases()
    scene = tmp_path / "gt"
    shutil.copytree(m.SCENE, scene)
    sample = dict(m.SAMPLE, path=str(scene))
    c = m.CONFIGS[tag]
    clf = m.make_clf(c)
    ld = dataLoader(clf, verbosity=0, device=DEV)
    ld.run(sample)
    assert not ld.cache_hit
    assert ld.getInfo() == int(gold[tag + "_n_nodes"])
    assert clf.temp.num_node_features == int(gold[tag + "_num_node_features"])
    assert ld.node_feature_names == list(gold[tag + "_node_names"])
    assert ld.features.dtype == torch.float32 and ld.features.device.type == "cpu"
    _ulp_close(ld.features.numpy(), gold[tag + "_features"], tag + " features")
    assert np.array_equal(ld.edge_lists.numpy(), gold[tag + "_edge_lists"]) and ld.edge_lists.dtype == torch.int64
    assert np.array_equal(ld.gt.numpy(), gold[tag + "_gt"]) and np.array_equal(ld.infinite.numpy(), gold[tag + "_infinite"])
    assert abs(ld.mean_edge - float(gold[tag + "_mean_edge"])) < 1e-15
    if c["edge_convs"]:
        assert ld.edge_feature_names == list(gold[tag + "_edge_names"])
        assert clf.temp.num_edge_features == int(gold[tag + "_num_edge_features"])
        _ulp_close(ld.edge_features.numpy(), gold[tag + "_edge_features"], tag + " edge features")
    else:
        assert clf.temp.num_edge_features is None and tuple(ld.edge_features.shape) == (1, 1)
    first = (ld.features.clone(), ld.edge_features.clone(), ld.edge_lists.clone(), ld.gt.clone(), ld.infinite.clone())
    ld2 = dataLoader(m.make_clf(c), verbosity=0, device=DEV, keep_on_device=True)
    ld2.run(sample)
    assert ld2.cache_hit and ld2.features.is_cuda
    assert torch.equal(ld2.features.cpu(), first[0]) and torch.equal(ld2.edge_lists, first[2])
    assert torch.equal(ld2.gt, first[3]) and torch.equal(ld2.infinite, first[4])
    if c["edge_convs"]:
        assert torch.equal(ld2.edge_features.cpu(), first[1]) and ld2.edge_feature_names == ld.edge_feature_names
    assert ld2.node_feature_names == ld.node_feature_names
    # a changed source file invalidates the cache
    import os, time
    p = str(scene / "7_cgeom.npz")
    z = dict(np.load(p)); z["radius"] = z["radius"] * 2.0
    np.savez(p, **z)
    os.utime(p, ns=(time.time_ns(), time.time_ns() + 10_000_000))
    ld3 = dataLoader(m.make_clf(c), verbosity=0, device=DEV)
    ld3.run(sample)
    assert not ld3.cache_hit


def test_loader_feeds_the_model_and_exports_scores(tmp_path, kf96_state):
    """Loader -> Data -> SurfaceNet.inference_layer -> exportScore, the chain of run.py:200-232 / data.py:521-535."""
    import shutil, os
    from dgnn_b200.data import dataLoader
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from oracle.static_model import SurfaceNet as OracleNet, make_clf, to_attr
    m, gold = _cases()
    scene = tmp_path / "gt"
    shutil.copytree(m.SCENE, scene)
    clf = m.make_clf(m.CONFIGS["kf96"])
    ld = dataLoader(clf, verbosity=0, device=DEV, keep_on_device=True)
    ld.run(dict(m.SAMPLE, path=str(scene)))
    ld.getInfo()
    mclf = make_clf(device=DEV, n_node_feat=clf.temp.num_node_features, n_edge_feat=clf.temp.num_edge_features)
    net = SurfaceNet(mclf); net.load_state_dict(kf96_state); net.to(DEV).eval()
    d = to_attr(dict(x=ld.features, edge_attr=ld.edge_features, edge_index=ld.edge_lists, y=ld.gt))
    z = net.inference_layer(d)
    ref = OracleNet(make_clf()); ref.load_state_dict(kf96_state); ref.eval()
    with torch.no_grad():
        zr = ref.inference_layer(to_attr(dict(x=torch.from_numpy(gold["kf96_features"]), edge_attr=torch.from_numpy(gold["kf96_edge_features"]),
                                              edge_index=torch.from_numpy(gold["kf96_edge_lists"]))))
    scale = np.maximum(np.abs(zr.numpy()), np.abs(zr.numpy()).mean())
    assert (np.abs(z.cpu().numpy() - zr.numpy()) / scale).max() <= 1e-4
    clf.paths.out = str(tmp_path)
    os.makedirs(tmp_path / "prediction")
    ld.exportScore(z.cpu())
    out = np.load(tmp_path / "prediction" / "7.npz")
    assert int(out["number_of_cells"]) == z.shape[0]
    np.testing.assert_allclose(out["softmax"], torch.softmax(z.cpu(), -1).numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["sigmoid"], torch.sigmoid(z.cpu()).numpy(), rtol=1e-5, atol=1e-7)
    assert np.array_equal(out["logits"], z.cpu().numpy())
