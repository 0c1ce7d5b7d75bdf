"""NumPy restatement of the loader's per-graph standardisation (TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``).

``processing/data.py:467-506``: ``StandardScaler().fit(X); X = scaler.transform(X)`` on the node features (columns 1..
when ``regularization.cell_type`` names column 0) and on the edge features.  sklearn 1.0.1 semantics restated: mean and
population variance per column accumulated in float64, ``scale_ = sqrt(var_)`` with (near-)constant columns set to 1
(``_handle_zeros_in_scale``), output in the input's float32.  ``tests/test_loader_cpu.py`` pins it to the installed
sklearn.
"""
import numpy as np


def standardize(x: np.ndarray, skip_first: bool = False) -> np.ndarray:
    out = np.array(x, dtype=np.float32, copy=True)
    sub = out[:, 1:] if skip_first else out
    x64 = sub.astype(np.float64)
    n = x64.shape[0]
    mean = x64.sum(axis=0) / n
    d = x64 - mean
    var = ((d * d).sum(axis=0) - d.sum(axis=0) ** 2 / n) / n
    scale = np.sqrt(np.maximum(var, 0.0))
    const = var <= 10 * np.finfo(np.float64).eps * n * mean * mean
    scale[const | (scale == 0)] = 1.0
    sub[:] = ((x64 - mean) / scale).astype(np.float32)
    return out


# --------------------------------------------------------------------------------------------------------------------
# The reference's ``dataLoader.run`` (``processing/data.py:81-112,194-284,353-414,434-519``) restated with NumPy only:
# an ordered column table standing in for the pandas DataFrame, sklearn's three scalers written out.  Pinned to the
# reference's own loader through tests/golden/loader_golden.npz (tests/golden/make_loader_golden.py).


class _Table:
    """Ordered name -> float64 column mapping with the two DataFrame operations the reference uses."""

    def __init__(self):
        self.names, self.cols = [], {}

    def assign(self, npz):
        for k in npz.files:
            if k not in self.cols:
                self.names.append(k)
            self.cols[k] = np.asarray(npz[k], dtype=np.float64)

    def drop(self, labels):
        for l in labels:
            if l not in self.cols:
                raise KeyError(l)
            self.names.remove(l)
            del self.cols[l]

    def insert_front(self, name, values):
        self.names.insert(0, name)
        self.cols[name] = np.array(values, dtype=np.float64, copy=True)

    def matrix(self):
        return np.stack([self.cols[k] for k in self.names], axis=1)


def _scale(x64, scaling, rng):
    """sklearn StandardScaler / MinMaxScaler / RobustScaler ``fit_transform`` on a float64 matrix."""
    eps = 10 * np.finfo(np.float64).eps
    n = x64.shape[0]
    if "s" in scaling:
        mean = x64.sum(axis=0) / n
        d = x64 - mean
        var = ((d * d).sum(axis=0) - d.sum(axis=0) ** 2 / n) / n
        scale = np.sqrt(np.maximum(var, 0.0))
        scale[(var <= eps * n * mean * mean) | (scale == 0)] = 1.0
        return (x64 - mean) / scale
    if "n" in scaling:
        lo, hi = rng
        dmin, dmax = x64.min(axis=0), x64.max(axis=0)
        r = dmax - dmin
        r[r < eps] = 1.0
        sc = (hi - lo) / r
        return x64 * sc + (lo - dmin * sc)
    if "r" in scaling:
        q25, q50, q75 = np.percentile(x64, [25, 50, 75], axis=0)
        iqr = q75 - q25
        iqr[iqr < eps] = 1.0
        return (x64 - q50) / iqr
    return x64


def load_graph(base, clf):
    """``dict(features f32[N,C], edge_features f32[4N,Ce] or None, edge_lists i64[2,4N], gt, infinite, node_names,
    edge_names, mean_edge)`` for the files ``base + "_*.npz"``."""
    f = clf.features
    stat = ("count", "dist_min", "dist_max", "dist_sum")
    sel = ("count", "min", "max", "sum")
    lab = np.load(base + "_labels.npz")
    if clf.inference.has_label:
        gt = np.stack([lab["inside_perc"], lab["outside_perc"]], axis=1).astype(np.float32)
    else:
        gt = np.zeros(lab["infinite"].shape, dtype=np.float32)
    infinite = lab["infinite"].astype(bool)
    t = _Table()
    geom = np.load(base + "_cgeom.npz")
    mean_edge = (geom["longest_edge"].sum() + geom["shortest_edge"].sum()) / (2 * len(geom["longest_edge"]))
    nf = f.node_features
    if "shape" in nf:
        t.assign(geom)
    if "vertex" in nf:
        t.assign(np.load(base + "_cbvf.npz"))
        for s, st in zip(sel, stat):
            if s not in nf:
                t.drop(["cb_vertex_%s_%s" % (w, st) for w in ("inside", "outside", "last")])
    if "facet" in nf:
        t.assign(np.load(base + "_cbff.npz"))
        if any(s not in nf for s in sel):
            raise AttributeError("NpzFile.drop (processing/data.py:245-253)")
    if "last" not in nf:
        if "vertex" in nf:
            t.drop(["cb_vertex_last_%s" % st for s, st in zip(sel, stat) if s in nf])
        if "facet" in nf:
            for s, st in zip(sel, stat):
                if s in nf:
                    t.drop(["cb_facet_last_first_%s" % st, "cb_facet_last_second_%s" % st])
    ct = clf.regularization.cell_type
    if ct:
        t.insert_front("reg_" + ct, t.cols[ct])
    adj = np.load(base + "_adjacencies.npz")["adjacencies"]
    edge_lists = adj.T.astype(np.int64)
    e = None
    if clf.model.edge_convs:
        e = _Table()
        ef = f.edge_features
        if "shape" in ef:
            e.assign(np.load(base + "_fgeom.npz"))
        for kind, suf in (("vertex", "_fbvf.npz"), ("facet", "_fbff.npz")):
            if kind in ef:
                e.assign(np.load(base + suf))
                if any(s not in ef for s in sel):
                    raise AttributeError("NpzFile.drop (processing/data.py:362-382)")
        if "last" not in ef:
            for kind in ("vertex", "facet"):
                if kind in ef:
                    e.drop(["fb_%s_last_%s" % (kind, st) for s, st in zip(sel, stat) if s in ef])
        et = clf.regularization.edge_type
        if et:
            e.insert_front("reg_" + et, e.cols[et])
    x = t.matrix()
    ex = e.matrix() if e is not None else None
    sc = f.scaling
    if sc:
        if "sum" in sc:
            if f.node_normalization_feature:
                x[:, 1:] = x[:, 1:] * 10 ** 3 / x[:, 1:].sum(axis=0)
            else:
                x = x * 10 ** 3 / x.sum(axis=0)
            if ex is not None:
                ex = ex * 10 ** 3 / ex.sum(axis=0)
        if "edge" in sc:
            x = x / mean_edge
        c0 = 1 if ct is not None else 0
        x[:, c0:] = _scale(x[:, c0:], sc, tuple(f.normalization_range))
        if ex is not None:
            c0 = 1 if clf.regularization.edge_type is not None else 0
            ex[:, c0:] = _scale(ex[:, c0:], sc, tuple(f.normalization_range))
    return dict(features=x.astype(np.float32), edge_features=ex.astype(np.float32) if ex is not None else None,
                edge_lists=edge_lists, gt=gt, infinite=infinite, node_names=list(t.names),
                edge_names=list(e.names) if e is not None else None, mean_edge=float(mean_edge))
