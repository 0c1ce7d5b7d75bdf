"""The loader step right before the model, on the device.

``processing/data.py`` standardises every graph's node and edge features with a fresh sklearn ``StandardScaler``
(``standardizeFeatures``, ``data.py:467-506``: mean / population standard deviation per column in float64, constant
columns left unscaled, first column kept raw when it is the regularisation feature) and turns the adjacency file into
``edge_index`` (``readAdjacencies_bin``, ``data.py:434-439``).  ``standardize_`` does the scaler's fit + transform with
two reductions and one elementwise kernel per matrix; the statistics are float64 partial sums reduced in a fixed order.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, lib, ptr


def edge_index_from_adjacencies(adjacencies) -> torch.Tensor:
    """``readAdjacencies_bin`` (``data.py:434-439``): ``adjacencies int32[4N,2]`` (row ``4i+k`` = (cell i, its k-th facet
    neighbour)) -> ``edge_index int64[2,4N]`` with ``[0]`` = owning cell (message source), ``[1]`` = neighbour (target)."""
    a = torch.as_tensor(np.asarray(adjacencies))
    return a.t().contiguous().to(torch.int64)


def standardize_(x: torch.Tensor, skip_first: bool = False) -> torch.Tensor:
    """In-place ``StandardScaler().fit_transform`` of a float32 ``[n, c]`` device matrix (columns 1.. when
    ``skip_first``: column 0 is the raw regularisation feature, ``data.py:485-488,501-506``).  Returns ``x``."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or not x.is_contiguous():
        raise ValueError("standardize_ expects a contiguous float32 CUDA matrix")
    n, ld = x.shape
    col0 = 1 if skip_first else 0
    st = torch.cuda.current_stream().cuda_stream
    blocks = lib().dgnn_small_grid()
    for c0 in range(col0, ld, 64):
        c = min(64, ld - c0)
        part = torch.empty((blocks, 2, c), dtype=torch.float64, device=x.device)
        call("dgnn_column_moments", ptr(x), n, ld, c0, c, None, ptr(part), blocks, st)
        mean = part[:, 0].sum(0) / n
        call("dgnn_column_moments", ptr(x), n, ld, c0, c, ptr(mean), ptr(part), blocks, st)
        s = part.sum(0)
        var = (s[1] - s[0] * s[0] / n) / n               # corrected two-pass variance (population)
        scale = var.clamp_min(0).sqrt()
        # sklearn _handle_zeros_in_scale: (near-)constant columns keep their values minus the mean
        const = var <= 10 * torch.finfo(torch.float64).eps * n * mean * mean
        scale = torch.where(const | (scale == 0), torch.ones_like(scale), scale)
        call("dgnn_column_affine", ptr(x), n, ld, c0, c, ptr(mean.contiguous()), ptr((1.0 / scale).contiguous()), ld,
             ptr(x), st)
    return x


# ------------------------------------------------------------------------------------------------------------------
# Loader front end: the reference's ``dataLoader`` (``processing/data.py:11-535``) with the same
# interface - ``run(d)``, ``getInfo()``, ``exportScore(prediction)``, the attributes ``features``, ``edge_features``,
# ``edge_lists``, ``gt``, ``infinite``, ``node_feature_names``, ``edge_feature_names``, ``mean_edge``, ``n_nodes`` - and the
# same column-order contract (npz key order = DataFrame column order = weight-matrix column order).  What changes: no
# pandas / sklearn; the float64 columns are uploaded once and scaled on the device in float64 (the StandardScaler path
# through dgnn_column_moments_f64 / dgnn_column_standardize_f64), rounded to float32 once like the reference's
# ``toTorch``; and the finished tensors are kept in a binary cache next to the source files so that the next run skips
# npz inflation, column assembly and scaling altogether.

_STAT = ("count", "dist_min", "dist_max", "dist_sum")
_SEL = ("count", "min", "max", "sum")
CACHE_VERSION = 2


def _in(key, sel):
    """The reference's ``'x' in clf.features.<list or str>`` (substring semantics when the value is a string)."""
    return sel is not None and key in sel


def _assemble(base, sel, spec, owner):
    """Column names and float64 arrays in the reference's order (``readNodeData_bin`` / ``readEdgeData_bin``).
    ``spec``: (shape file, vertex file, vertex prefix, facet file, facet last-column stems)."""
    import numpy as np
    shape_f, vert_f, vert_p, facet_f, facet_last = spec
    names, cols = [], []

    def add(npz):
        for k in npz.files:
            if k in names:                       # DataFrame assignment to an existing column overwrites in place
                cols[names.index(k)] = np.asarray(npz[k], dtype=np.float64)
            else:
                names.append(k); cols.append(np.asarray(npz[k], dtype=np.float64))

    def drop(labels):
        for l in labels:
            if l not in names:
                raise KeyError("%r not found in axis" % [l])            # pandas.DataFrame.drop
            i = names.index(l)
            del names[i]; del cols[i]

    extra = {}
    if owner == "node":
        geom = np.load(base + shape_f)
        extra["mean_edge"] = (geom["longest_edge"].sum() + geom["shortest_edge"].sum()) / (2 * len(geom["longest_edge"]))
        if _in("shape", sel):
            add(geom)
    elif _in("shape", sel):
        add(np.load(base + shape_f))
    if _in("vertex", sel):
        add(np.load(base + vert_f))
        for s, st in zip(_SEL, _STAT):
            if not _in(s, sel):
                if owner == "edge":   # data.py:362-369 calls .drop on the NpzFile itself
                    raise AttributeError("'NpzFile' object has no attribute 'drop' (processing/data.py:362-369: dropping "
                                         "a statistic from the edge features is broken in the reference)")
                drop(["%s_%s_%s" % (vert_p, w, st) for w in ("inside", "outside", "last")])
    if _in("facet", sel):
        add(np.load(base + facet_f))
        for s in _SEL:
            if not _in(s, sel):       # data.py:245-253 / :375-382 call .drop on the NpzFile itself
                raise AttributeError("'NpzFile' object has no attribute 'drop' (processing/data.py:245-253,375-382: "
                                     "dropping a statistic from the facet features is broken in the reference)")
    if not _in("last", sel):
        if _in("vertex", sel):
            drop(["%s_last_%s" % (vert_p, st) for s, st in zip(_SEL, _STAT) if _in(s, sel)])
        if _in("facet", sel):
            for s, st in zip(_SEL, _STAT):
                if _in(s, sel):
                    drop([stem % st for stem in facet_last])
    return names, cols, extra


class dataLoader:
    """``processing/data.py:dataLoader`` on the device.  ``device``: where the scaling runs (and, with
    ``keep_on_device=True``, where ``features`` / ``edge_features`` stay - ``SurfaceNet`` takes them as they are);
    ``cache``: keep / use the binary cache ``<basefilename>_dgnn_<key>.bin``."""

    NODE_SPEC = ("_cgeom.npz", "_cbvf.npz", "cb_vertex", "_cbff.npz", ("cb_facet_last_first_%s", "cb_facet_last_second_%s"))
    EDGE_SPEC = ("_fgeom.npz", "_fbvf.npz", "fb_vertex", "_fbff.npz", ("fb_facet_last_%s",))

    def __init__(self, clf, verbosity=1, device="cuda:0", keep_on_device=False, cache=True):
        self.n_nodes = 0
        self.id = ''
        self.gt = []
        self.features = []
        self.edge_features = []
        self.edge_lists = []
        self.clf = clf
        self.verbosity = verbosity
        self.read_edge_features = clf.model.edge_convs
        self.device = torch.device(device)
        self.keep_on_device = keep_on_device
        self.cache = cache
        self.cache_hit = False

    def __len__(self):
        return len(self.gt)

    def reset(self):
        self.n_nodes = 0
        self.gt = []
        self.features = []
        self.edge_features = []
        self.edge_lists = []
        self.id = ''

    def getInfo(self):   # data.py:39-79
        fs = self.features.size()[1] - bool(self.clf.regularization.cell_type)
        self.clf.temp.num_node_features = fs
        if self.read_edge_features:
            fse = self.edge_features.size()[1] - bool(self.clf.regularization.edge_type)
            self.clf.temp.num_edge_features = fse
        else:
            self.clf.temp.num_edge_features = None
        if self.verbosity:
            print("\t-{} nodes".format(self.n_nodes))
            print("\t-{} node features:".format(fs))
            print("\t-", self.node_feature_names[1:] if self.clf.features.node_normalization_feature else self.node_feature_names)
            if self.read_edge_features:
                print("\t-{} edge features:".format(fse))
                print("\t-", self.edge_feature_names[1:] if self.clf.features.edge_normalization_feature else self.edge_feature_names)
        return self.n_nodes

    # ------------------------------------------------------------------ run
    def run(self, d):   # data.py:81-112
        import os
        self.path, self.filename, self.category = d["path"], d["filename"], d["category"]
        self.id, self.scan_conf, self.gtfile, self.ioufile = d["id"], d["scan_conf"], d["gtfile"], d["ioufile"]
        self.basefilename = os.path.join(self.path, self.gtfile)
        self.cache_hit = False
        if self.cache and self._cache_load():
            self.cache_hit = True
        else:
            self.readNodeData_bin()
            self.readAdjacencies_bin()
            if self.read_edge_features:
                self.readEdgeData_bin()
            if self.clf.features.scaling:
                self.standardizeFeatures()
            self.toTorch()
            if self.cache:
                self._cache_store()
        self.n_nodes += len(self.features)

    def readNodeData_bin(self):   # data.py:194-284
        import numpy as np
        temp = np.load(self.basefilename + "_labels.npz")
        if self.clf.inference.has_label:
            self.gt = torch.from_numpy(np.stack([temp["inside_perc"], temp["outside_perc"]], axis=1)).to(torch.float)
        else:
            self.gt = torch.zeros(size=temp["infinite"].shape)
        self.infinite = torch.from_numpy(temp["infinite"]).type(torch.bool)
        names, cols, extra = _assemble(self.basefilename, self.clf.features.node_features, self.NODE_SPEC, "node")
        self.mean_edge = extra["mean_edge"]
        ct = self.clf.regularization.cell_type
        if ct:
            names.insert(0, "reg_" + ct); cols.insert(0, cols[names.index(ct, 1) - 1].copy())
        self.node_feature_names = names
        self._node_cols = np.stack(cols, axis=1) if cols else np.zeros((self.gt.shape[0], 0))
        assert self.gt.shape[0] == self._node_cols.shape[0]
        assert not np.isnan(self._node_cols).any()
        assert not torch.isnan(self.gt).any()

    def readAdjacencies_bin(self):   # data.py:434-439
        import numpy as np
        temp = np.load(self.basefilename + "_adjacencies.npz")
        self.edge_lists = edge_index_from_adjacencies(temp["adjacencies"])

    def readEdgeData_bin(self):   # data.py:353-414
        import numpy as np
        names, cols, _ = _assemble(self.basefilename, self.clf.features.edge_features, self.EDGE_SPEC, "edge")
        et = self.clf.regularization.edge_type
        if et:
            names.insert(0, "reg_" + et); cols.insert(0, cols[names.index(et, 1) - 1].copy())
        self.edge_feature_names = names
        self._edge_cols = np.stack(cols, axis=1)
        assert self.edge_lists.shape[1] == self._edge_cols.shape[0]
        assert not np.isnan(self._edge_cols).any()

    # ------------------------------------------------------------------ scaling (data.py:444-506), float64 on the device
    def _fit_transform(self, x64, col0, scaling):
        """x64: float64 CUDA matrix; returns the float32 matrix with columns col0.. scaled, the others cast."""
        n, c = x64.shape
        out = x64.to(torch.float32)
        if c - col0 <= 0 or n == 0:
            return out
        if _in('s', scaling):                                  # StandardScaler
            st = torch.cuda.current_stream().cuda_stream
            blocks = lib().dgnn_small_grid()
            for c0 in range(col0, c, 64):
                cc = min(64, c - c0)
                part = torch.empty((blocks, 2, cc), dtype=torch.float64, device=x64.device)
                call("dgnn_column_moments_f64", ptr(x64), n, c, c0, cc, None, ptr(part), blocks, st)
                mean = (part[:, 0].sum(0) / n).contiguous()
                call("dgnn_column_moments_f64", ptr(x64), n, c, c0, cc, ptr(mean), ptr(part), blocks, st)
                s = part.sum(0)
                var = (s[1] - s[0] * s[0] / n) / n
                scale = var.clamp_min(0).sqrt()
                const = var <= 10 * torch.finfo(torch.float64).eps * n * mean * mean     # sklearn _handle_zeros_in_scale
                scale = torch.where(const | (scale == 0), torch.ones_like(scale), scale).contiguous()
                call("dgnn_column_standardize_f64", ptr(x64), n, c, c0, cc, ptr(mean), ptr(scale), c, c0, ptr(out), st)
            return out
        sub = x64[:, col0:]
        eps = 10 * torch.finfo(torch.float64).eps
        if _in('n', scaling):                                  # MinMaxScaler(feature_range)
            lo, hi = tuple(self.clf.features.normalization_range)
            dmin, dmax = sub.min(0).values, sub.max(0).values
            rng = dmax - dmin
            rng = torch.where(rng < eps, torch.ones_like(rng), rng)
            sc = (hi - lo) / rng
            out[:, col0:] = (sub * sc + (lo - dmin * sc)).to(torch.float32)
        elif _in('r', scaling):                                # RobustScaler: median / inter-quartile range
            q = torch.quantile(sub, torch.tensor([0.25, 0.5, 0.75], dtype=torch.float64, device=sub.device), dim=0)
            iqr = q[2] - q[0]
            iqr = torch.where(iqr < eps, torch.ones_like(iqr), iqr)
            out[:, col0:] = ((sub - q[1]) / iqr).to(torch.float32)
        return out

    def standardizeFeatures(self):
        import numpy as np
        f = self.clf.features
        sc = f.scaling
        dev = self.device
        x = torch.from_numpy(self._node_cols).to(dev)
        e = torch.from_numpy(self._edge_cols).to(dev) if self.read_edge_features else None
        names = self.node_feature_names
        ct, et = self.clf.regularization.cell_type, self.clf.regularization.edge_type
        if _in('sum', sc):
            if f.node_normalization_feature:
                x[:, 1:] = x[:, 1:] * 10 ** 3 / x[:, 1:].sum(0)
            else:
                x = x * 10 ** 3 / x.sum(0)
            if e is not None:
                e = e * 10 ** 3 / e.sum(0)
        if _in('vol', sc):
            raise NotImplementedError("scaling 'vol' reads DataFrame.norm, which does not exist (processing/data.py:455-460)")
        if f.node_normalization_feature is not None:
            if ct is None:
                raise KeyError(None)                               # features[None] in the reference
            x[:, 1:] = x[:, 1:] / (x[:, names.index(ct, 1)] + 0.0001)[:, None]
        if _in('edge', sc):
            x = x / self.mean_edge
        known = _in('s', sc) or _in('n', sc) or _in('r', sc)
        if not known and not _in('sum', sc):
            raise ValueError("{} are no valid scalers. choose either 'sum', 's', 'n' or 'r'".format(sc))
        if not known:
            self._x32, self._e32 = x.to(torch.float32), (e.to(torch.float32) if e is not None else None)
            return
        self._x32 = self._fit_transform(x.contiguous(), 1 if ct is not None else 0, sc)
        self._e32 = None
        if e is not None:
            if f.edge_normalization_feature is not None:
                if et is None:
                    raise KeyError(None)
                e[:, 1:] = e[:, 1:] / (e[:, self.edge_feature_names.index(et, 1)] + 0.0001)[:, None]
            self._e32 = self._fit_transform(e.contiguous(), 1 if et is not None else 0, sc)

    def toTorch(self):   # data.py:512-519
        x32 = getattr(self, "_x32", None)
        if x32 is None:                                            # no scaling configured: plain cast
            x32 = torch.from_numpy(self._node_cols).to(torch.float32)
            e32 = torch.from_numpy(self._edge_cols).to(torch.float32) if self.read_edge_features else None
        else:
            e32 = self._e32
        keep = self.keep_on_device
        self.features = x32 if keep else x32.cpu()
        if self.read_edge_features:
            self.edge_features = e32 if keep else e32.cpu()
        else:
            self.edge_features = torch.empty(1, 1, dtype=torch.float)
        self._x32 = self._e32 = self._node_cols = self._edge_cols = None

    # ------------------------------------------------------------------ binary cache
    def _cache_key(self):
        import hashlib, json, os
        f = self.clf.features
        src = []
        for suf in ("_labels.npz", "_cgeom.npz", "_cbvf.npz", "_cbff.npz", "_adjacencies.npz", "_fgeom.npz", "_fbvf.npz", "_fbff.npz"):
            p = self.basefilename + suf
            if os.path.exists(p):
                s = os.stat(p)
                src.append((suf, s.st_size, s.st_mtime_ns))
        cfg = dict(v=CACHE_VERSION, scaling=f.scaling, rng=list(f.normalization_range) if f.get("normalization_range") else None,
                   node=f.node_features, edge=f.edge_features, nnf=f.node_normalization_feature,
                   enf=f.edge_normalization_feature, ct=self.clf.regularization.cell_type, et=self.clf.regularization.edge_type,
                   edges=bool(self.read_edge_features), has_label=bool(self.clf.inference.has_label), src=src)
        return hashlib.sha1(json.dumps(cfg, sort_keys=True, default=str).encode()).hexdigest()[:16]

    def _cache_path(self):
        return "%s_dgnn_%s.bin" % (self.basefilename, self._cache_key())

    def _cache_store(self):
        import json, os
        import numpy as np
        arrays = {"features": self.features, "gt": self.gt, "infinite": self.infinite.to(torch.uint8),
                  "adjacencies": self.edge_lists.t().to(torch.int32).contiguous()}
        if self.read_edge_features:
            arrays["edge_features"] = self.edge_features
        meta, blobs, off = {}, [], 0
        for k, t in arrays.items():
            a = np.ascontiguousarray(t.detach().cpu().numpy())
            meta[k] = dict(dtype=str(a.dtype), shape=list(a.shape), offset=off)
            blobs.append(a)
            off += (a.nbytes + 63) // 64 * 64
        head = json.dumps(dict(version=CACHE_VERSION, arrays=meta, node_feature_names=self.node_feature_names,
                               edge_feature_names=getattr(self, "edge_feature_names", None),
                               mean_edge=float(self.mean_edge))).encode()
        path = self._cache_path()
        tmp = path + ".tmp%d" % os.getpid()
        try:
            with open(tmp, "wb") as fh:
                fh.write(b"DGNNBIN1" + np.uint64(len(head)).tobytes() + head)
                pad = (-fh.tell()) % 64
                fh.write(b"\0" * pad)
                for a in blobs:
                    fh.write(a.tobytes())
                    fh.write(b"\0" * ((-a.nbytes) % 64))
            os.replace(tmp, path)
        except OSError:                                            # read-only data directory: run without the cache
            if os.path.exists(tmp):
                os.remove(tmp)

    def _cache_load(self):
        import json, os
        import numpy as np
        path = self._cache_path()
        if not os.path.exists(path):
            return False
        with open(path, "rb") as fh:
            if fh.read(8) != b"DGNNBIN1":
                return False
            n = int(np.frombuffer(fh.read(8), dtype=np.uint64)[0])
            head = json.loads(fh.read(n).decode())
            base = (16 + n + 63) // 64 * 64
        if head.get("version") != CACHE_VERSION:
            return False
        mm = np.memmap(path, dtype=np.uint8, mode="r")

        def arr(k):
            m = head["arrays"][k]
            cnt = int(np.prod(m["shape"])) if m["shape"] else 1
            a = np.frombuffer(mm, dtype=np.dtype(m["dtype"]), count=cnt, offset=base + m["offset"]).reshape(m["shape"])
            return torch.from_numpy(np.array(a))                   # one copy out of the page cache

        dev = self.device if self.keep_on_device else None
        self.features = arr("features").to(dev) if dev else arr("features")
        self.gt = arr("gt")
        self.infinite = arr("infinite").to(torch.bool)
        self.edge_lists = edge_index_from_adjacencies(arr("adjacencies").numpy())
        if self.read_edge_features:
            self.edge_features = arr("edge_features").to(dev) if dev else arr("edge_features")
            self.edge_feature_names = head["edge_feature_names"]
        else:
            self.edge_features = torch.empty(1, 1, dtype=torch.float)
        self.node_feature_names = head["node_feature_names"]
        self.mean_edge = head["mean_edge"]
        return True

    def exportScore(self, prediction):   # data.py:521-535
        import os
        import numpy as np
        from .runModel import export_scores
        outpath = os.path.join(self.clf.paths.out, "prediction")
        if self.verbosity:
            print("Export predictions to: ", outpath)
        file = os.path.join(outpath, self.filename + ".npz")
        # the reference hands over CPU logits (runModel.py:451); the scores are computed on the device either way
        sc = export_scores(prediction if prediction.is_cuda else prediction.to(self.device))
        with open(file, 'wb') as f:
            np.savez(f, **sc)
