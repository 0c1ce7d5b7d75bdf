"""Synthetic inputs for benchmarks, smoke runs and tests, plus attribute-dict stand-ins for the
reference's ``Munch`` config objects.

The reference has no generator (its graphs come from the external CGAL tool ``feat``); the metric
is measured on synthetic Delaunay tetrahedralisations of the same shape:
scipy ``Delaunay`` of random or scan-like points, one infinite cell per hull facet, adjacency in the
reference's on-disk layout (``processing/data.py:434-439``), random features of the ``feat`` tool's
shape.  Host-side NumPy; nothing here is on the timed path.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- points


def random_points(n_points: int, seed: int = 0) -> np.ndarray:
    """Uniform random points in the unit cube ("random")."""
    rng = np.random.default_rng(seed)
    return rng.random((n_points, 3))


def scan_like_points(n_points: int, seed: int = 0, sigma: float = 0.005,
                     outliers: float = 0.02) -> np.ndarray:
    """Scan-like points: a closed surface (sphere + torus) with Gaussian noise and an
    outlier fraction, mirroring the scan confs of ``processing/modelnet/scan.py:10-44``."""
    rng = np.random.default_rng(seed)
    n_out = int(n_points * outliers)
    n_surf = n_points - n_out
    n_sph = n_surf // 2
    n_tor = n_surf - n_sph
    v = rng.standard_normal((n_sph, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    sph = 0.5 + 0.3 * v
    a = rng.random(n_tor) * 2 * np.pi
    b = rng.random(n_tor) * 2 * np.pi
    tor = np.stack([(0.35 + 0.08 * np.cos(b)) * np.cos(a),
                    (0.35 + 0.08 * np.cos(b)) * np.sin(a),
                    0.08 * np.sin(b)], axis=1) + 0.5
    pts = np.concatenate([sph, tor], axis=0)
    pts += rng.standard_normal(pts.shape) * sigma
    out = rng.random((n_out, 3))
    return np.concatenate([pts, out], axis=0)


# --------------------------------------------------------------------------- Delaunay graph


def delaunay_graph(points: np.ndarray):
    """3D Delaunay graph in the reference's file layout.

    Returns ``(adjacencies int32[4N,2], infinite int32[N], centroids float64[N,3],
    tetrahedra int32[T,4])``.

    Convention: finite cells keep scipy order and slot ``k`` is the
    neighbour opposite vertex ``k`` (the CGAL convention the reference data uses); hull
    facets enumerated in ``(tet, k)`` order of ``neighbors == -1`` become infinite cells
    ``T, T+1, ...``; infinite-cell slot 0 is its finite cell, slots 1-3 are the infinite
    cells across the hull facet's three edges ordered by the facet vertex opposite that
    edge (ascending vertex id); the finite cell's ``-1`` slot is patched with the
    infinite id.  Every node then has exactly 4 distinct neighbours, as in the real data.
    """
    from scipy.spatial import Delaunay

    tri = Delaunay(points)
    simp = tri.simplices.astype(np.int64)
    nbr = tri.neighbors.astype(np.int64).copy()
    T = simp.shape[0]
    ht, hk = np.nonzero(nbr == -1)  # row-major == (tet, k) order
    H = ht.shape[0]
    inf_id = T + np.arange(H, dtype=np.int64)
    nbr[ht, hk] = inf_id
    # hull facet vertices: the three vertices of tet ht except vertex hk, ascending id
    mask = np.ones((H, 4), dtype=bool)
    mask[np.arange(H), hk] = False
    fv = np.sort(simp[ht][mask].reshape(H, 3), axis=1)  # [H,3] ascending
    # edge opposite facet-vertex j is the pair of the other two vertices
    opp = [(1, 2), (0, 2), (0, 1)]
    V = int(points.shape[0])
    keys = np.empty((H, 3), dtype=np.int64)
    for j, (a, b) in enumerate(opp):
        keys[:, j] = fv[:, a] * V + fv[:, b]  # fv ascending => a<b
    flat = keys.reshape(-1)
    order = np.argsort(flat, kind="stable")
    sk = flat[order]
    # each hull edge is shared by exactly two hull facets
    assert sk.shape[0] % 2 == 0 and np.all(sk[0::2] == sk[1::2]), "hull is not a closed 2-manifold"
    partner = np.empty_like(order)
    partner[order[0::2]] = order[1::2]
    partner[order[1::2]] = order[0::2]
    inf_nbr = np.empty((H, 4), dtype=np.int64)
    inf_nbr[:, 0] = ht
    inf_nbr[:, 1:] = T + (partner.reshape(H, 3) // 3)
    full = np.concatenate([nbr, inf_nbr], axis=0)
    N = T + H
    adj = np.empty((4 * N, 2), dtype=np.int32)
    adj[:, 0] = np.repeat(np.arange(N, dtype=np.int32), 4)
    adj[:, 1] = full.reshape(-1).astype(np.int32)
    infinite = np.zeros(N, dtype=np.int32)
    infinite[T:] = 1
    cen = np.empty((N, 3), dtype=np.float64)
    cen[:T] = points[simp].mean(axis=1)
    cen[T:] = points[fv].mean(axis=1)
    return adj, infinite, cen, simp.astype(np.int32)


def lattice_graph(nx: int, ny: int, nz: int):
    """Analytic 4-regular periodic graph for the largest configs (SURVEY.md section 7
    "hard parts"): the diamond-cubic lattice (each site has 4 neighbours, like tetrahedra
    of a tetrahedralisation), periodic in all axes.  Returns ``(adjacencies, infinite,
    centroids)``; no infinite cells.  Site = (cell x,y,z, sublattice s in {0,1});
    sublattice-0 site (x,y,z) bonds to sublattice-1 sites at (x,y,z), (x-1,y,z),
    (x,y-1,z), (x,y,z-1); slot k of either end is the same bond direction, so the
    reverse slot of slot k is k."""
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    x = x.reshape(-1); y = y.reshape(-1); z = z.reshape(-1)
    C = nx * ny * nz

    def sid(xx, yy, zz, s):
        return (((xx % nx) * ny + (yy % ny)) * nz + (zz % nz)) * 2 + s

    N = 2 * C
    nbr = np.empty((N, 4), dtype=np.int64)
    a = sid(x, y, z, 0)
    b = sid(x, y, z, 1)
    nbr[a, 0] = sid(x, y, z, 1)
    nbr[a, 1] = sid(x - 1, y, z, 1)
    nbr[a, 2] = sid(x, y - 1, z, 1)
    nbr[a, 3] = sid(x, y, z - 1, 1)
    nbr[b, 0] = sid(x, y, z, 0)
    nbr[b, 1] = sid(x + 1, y, z, 0)
    nbr[b, 2] = sid(x, y + 1, z, 0)
    nbr[b, 3] = sid(x, y, z + 1, 0)
    adj = np.empty((4 * N, 2), dtype=np.int32)
    adj[:, 0] = np.repeat(np.arange(N, dtype=np.int32), 4)
    adj[:, 1] = nbr.reshape(-1).astype(np.int32)
    cen = np.empty((N, 3), dtype=np.float64)
    cen[a] = np.stack([x, y, z], axis=1) + 0.25
    cen[b] = np.stack([x, y, z], axis=1) + 0.75
    return adj, np.zeros(N, dtype=np.int32), cen


def synthetic_features(n_cells: int, infinite: np.ndarray, seed: int = 1,
                       n_node_feat: int = 28, n_edge_feat: int = 20):
    """Random features of the ``feat`` tool's shape.

    ``x = [w | f]`` with ``w`` the raw volume-like loss weight (0 for infinite cells),
    ``edge_attr`` independent per directed edge, ``y = (u, 1-u)``.
    """
    rng = np.random.default_rng(seed)
    w = rng.random(n_cells) * 1e-3
    w[infinite.astype(bool)] = 0.0
    f = rng.standard_normal((n_cells, n_node_feat))
    x = np.concatenate([w[:, None], f], axis=1).astype(np.float32)
    ea = rng.standard_normal((4 * n_cells, n_edge_feat)).astype(np.float32)
    u = rng.random(n_cells)
    y = np.stack([u, 1.0 - u], axis=1).astype(np.float32)
    return x, ea, y



# --------------------------------------------------------------------------- config stand-ins


class AttrDict(dict):
    """Attribute-access dict standing in for ``munch.Munch`` (``run.py:291``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, list):
        return [to_attr(v) for v in d]
    return d


def make_clf(convs=(64, 128, 128, 128), edge_convs=1, decoder=2, normalization='b', loss='kl',
             cell_type='vol', edge_type=None, cell_norm=None, n_node_feat=28, n_edge_feat=20,
             device='cpu'):
    """A ``clf`` with the keys the Static model / trainer read."""
    return to_attr(dict(
        model=dict(type='sage', convs=list(convs), edge_convs=edge_convs, decoder=decoder,
                   normalization=normalization, edge_prediction=0),
        training=dict(loss=loss, learning_rate=0.005),
        regularization=dict(cell_type=cell_type, edge_type=edge_type, cell_norm=cell_norm,
                            edge_epoch=None, edge_weight=0.4),
        graph=dict(num_hops=len(convs), additional_num_hops=1, self_loops=0),
        inference=dict(per_layer=1, has_label=1, batch_size=0),
        temp=dict(num_node_features=n_node_feat, num_edge_features=n_edge_feat, device=device,
                  batch_size=0, current_epoch=0),
    ))
