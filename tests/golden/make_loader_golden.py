"""Golden vectors for the loader front end: a small synthetic object in the reference's on-disk
format (tests/golden/loader_scene/gt/7_*.npz, key names and order of the real data/Ignatius files) is read by the
reference's UNMODIFIED ``processing/data.py`` (pandas + scikit-learn, runs in the build container) under three feature
configurations; its outputs are committed as tests/golden/loader_golden.npz.

    python tests/golden/make_loader_golden.py        # needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import graph as og                      # noqa: E402
from oracle.static_model import to_attr             # noqa: E402

SCENE = os.path.join(HERE, "loader_scene", "gt")
STAT = ("count", "dist_min", "dist_max", "dist_sum")

CONFIGS = {
    # configs/pretrained/reconbench.yaml: everything, standardised, raw volume in column 0
    "kf96": dict(scaling="s", node_features=["shape", "vertex", "facet", "count", "min", "max", "sum", "first", "second"],
                 edge_features=["shape", "vertex", "facet", "count", "min", "max", "sum"], cell_type="vol", edge_type=None,
                 has_label=1, edge_convs=1),
    # no regularisation column, 'last' statistics kept, min-max scaling
    "minmax_last": dict(scaling="n", node_features=["shape", "vertex", "facet", "count", "min", "max", "sum", "last"],
                        edge_features=["shape", "vertex", "facet", "count", "min", "max", "sum", "last"], cell_type=None,
                        edge_type=None, has_label=0, edge_convs=1),
    # a statistic dropped from the node vertex features, sum-scaling followed by the StandardScaler ('s' in 'sum')
    "sum_nocount": dict(scaling="sum", node_features=["shape", "vertex", "min", "max", "sum"],
                        edge_features=["shape", "vertex", "facet", "count", "min", "max", "sum"], cell_type="vol", edge_type=None,
                        has_label=1, edge_convs=1),
    # robust scaling, no edge features read
    "robust_noedge": dict(scaling="r", node_features=["shape", "vertex", "facet", "count", "min", "max", "sum"],
                          edge_features=[], cell_type="vol", edge_type=None, has_label=1, edge_convs=0),
}


def make_clf(c):
    return to_attr(dict(model=dict(edge_convs=c["edge_convs"]), inference=dict(has_label=c["has_label"]),
                        regularization=dict(cell_type=c["cell_type"], edge_type=c["edge_type"]),
                        features=dict(scaling=c["scaling"], normalization_range=[0, 1], node_features=c["node_features"],
                                      edge_features=c["edge_features"], node_normalization_feature=None,
                                      edge_normalization_feature=None),
                        temp=dict(), paths=dict(out="/tmp")))


SAMPLE = dict(path=SCENE, filename="7", category="synthetic", id="7", scan_conf=0, gtfile="7", ioufile="7")


def write_scene(seed=3, n_points=70):
    """Files of one object: sparse visibility statistics like the feat tool's (most entries at the column mode)."""
    rng = np.random.default_rng(seed)
    adj, infinite, cen, tets = og.delaunay_graph(og.scan_like_points(n_points, seed=seed))
    n = infinite.shape[0]
    os.makedirs(SCENE, exist_ok=True)

    def sparse(m, scale):
        v = rng.gamma(2.0, scale, m)
        v[rng.random(m) < 0.7] = 0.0
        return v

    np.savez(os.path.join(SCENE, "7_labels.npz"), infinite=infinite.astype(np.int32), inside_perc=rng.random(n),
             outside_perc=rng.random(n))
    vol = rng.gamma(2.0, 1e-4, n); vol[infinite.astype(bool)] = 0.0
    np.savez(os.path.join(SCENE, "7_cgeom.npz"), radius=rng.gamma(2.0, 0.05, n), vol=vol, longest_edge=rng.gamma(3.0, 0.05, n),
             shortest_edge=rng.gamma(2.0, 0.02, n))
    np.savez(os.path.join(SCENE, "7_cbvf.npz"), **{"cb_vertex_%s_%s" % (w, s): (np.floor(sparse(n, 2.0)) if s == "count" else sparse(n, 0.3))
                                                   for w in ("inside", "outside", "last") for s in STAT})
    np.savez(os.path.join(SCENE, "7_cbff.npz"), **{"cb_facet_%s_%s_%s" % (w, o, s): (np.floor(sparse(n, 2.0)) if s == "count" else sparse(n, 0.3))
                                                   for w in ("inside", "outside", "last") for o in ("first", "second") for s in STAT})
    np.savez(os.path.join(SCENE, "7_adjacencies.npz"), adjacencies=adj.astype(np.int32))
    e = 4 * n
    fg = {"area": rng.gamma(2.0, 0.01, e), "angle": rng.random(e) * np.pi, "cc": rng.gamma(2.0, 0.05, e), "beta": rng.random(e)}
    fg["cc"][rng.random(e) < 0.05] = 7.5      # repeated values
    np.savez(os.path.join(SCENE, "7_fgeom.npz"), **fg)
    np.savez(os.path.join(SCENE, "7_fbvf.npz"), **{"fb_vertex_%s_%s" % (w, s): (np.floor(sparse(e, 2.0)) if s == "count" else sparse(e, 0.3))
                                                   for w in ("inside", "outside", "last") for s in STAT})
    np.savez(os.path.join(SCENE, "7_fbff.npz"), **{"fb_facet_%s_%s" % (w, s): (np.floor(sparse(e, 2.0)) if s == "count" else sparse(e, 0.3))
                                                   for w in ("inside", "outside", "last") for s in STAT})
    # one constant column (zero variance: the scalers leave it unscaled)
    z = dict(np.load(os.path.join(SCENE, "7_fbff.npz")))
    z["fb_facet_outside_dist_max"][:] = 0.25
    np.savez(os.path.join(SCENE, "7_fbff.npz"), **z)


def main():
    write_scene()
    sys.path.insert(0, "/root/reference")
    from processing.data import dataLoader          # the reference's own loader, unmodified
    out = {}
    for tag, c in CONFIGS.items():
        clf = make_clf(c)
        ld = dataLoader(clf, verbosity=0)
        ld.run(SAMPLE)
        n = ld.getInfo()
        out[tag + "_features"] = ld.features.numpy()
        out[tag + "_gt"] = ld.gt.numpy()
        out[tag + "_infinite"] = ld.infinite.numpy()
        out[tag + "_edge_lists"] = ld.edge_lists.numpy()
        out[tag + "_node_names"] = np.array(ld.node_feature_names)
        out[tag + "_n_nodes"] = np.int64(n)
        out[tag + "_num_node_features"] = np.int64(clf.temp.num_node_features)
        out[tag + "_mean_edge"] = np.float64(ld.mean_edge)
        if c["edge_convs"]:
            out[tag + "_edge_features"] = ld.edge_features.numpy()
            out[tag + "_edge_names"] = np.array(ld.edge_feature_names)
            out[tag + "_num_edge_features"] = np.int64(clf.temp.num_edge_features)
    np.savez_compressed(os.path.join(HERE, "loader_golden.npz"), **out)
    print("wrote", len(out), "arrays;", {k: v.shape for k, v in out.items() if k.endswith("features")})


if __name__ == "__main__":
    main()
