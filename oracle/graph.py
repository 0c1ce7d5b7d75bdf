"""NumPy oracle for the graph side of the path (TEST INFRASTRUCTURE ONLY).

Restates / defines, on the CPU with plain NumPy:

* the on-disk adjacency layout of the reference loader
  (``processing/data.py:434-439``: ``adjacencies int32[4N,2]``, row ``4i+k`` =
  ``(i, k-th facet neighbour of i)``; ``edge_index = adjacencies.T``);
* the synthetic Delaunay graph convention of SURVEY.md Appendix D (the reference has
  no generator; its inputs come from the external ``feat`` tool);
* the ELL-4 table, reverse-facet slot, Morton permutation, edge re-layout and the
  partition / halo maps that the CUDA/C++ builders must reproduce bit-exactly.
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


# --------------------------------------------------------------------------- ELL-4 layout


def ell_from_adjacency(adj: np.ndarray):
    """ELL-4 table + reverse-facet slot from the reference adjacency layout.

    ``nbr[i,k] = adjacencies[4i+k,1]`` (``processing/data.py:434-439``);
    ``rslot[i,k] = k'`` such that ``nbr[nbr[i,k], k'] == i`` (first match).  The edge row
    that carries the message INTO ``i`` through slot ``k`` is then
    ``4*nbr[i,k] + rslot[i,k]`` (source = edge_index[0] = owning cell,
    ``surfaceNetStaticEdgeFilters.py:80``).
    """
    N = adj.shape[0] // 4
    assert adj.shape == (4 * N, 2)
    assert np.array_equal(adj[:, 0], np.repeat(np.arange(N, dtype=adj.dtype), 4)), \
        "adjacency rows must be grouped 4 per owning cell"
    nbr = adj[:, 1].reshape(N, 4).astype(np.int32)
    back = nbr[nbr]  # [N,4,4]: neighbours of my neighbours
    hit = back == np.arange(N, dtype=np.int32)[:, None, None]
    assert hit.any(axis=2).all(), "adjacency is not symmetric"
    rslot = hit.argmax(axis=2).astype(np.uint8)
    return nbr, rslot


def ell_from_edges(src: np.ndarray, tgt: np.ndarray, n_rows: int):
    """Generic ELL-4 build from an edge list: for every row (target) the <=4 incident
    edges in ascending edge id.  Returns ``(nbr int32[n_rows,4], eid int32[n_rows,4],
    cnt int32[n_rows])`` with -1 padding.  This is what PyG's scatter-mean over
    ``edge_index[1]`` sees (``surfaceNetStaticEdgeFilters.py:80``)."""
    E = src.shape[0]
    order = np.argsort(tgt, kind="stable")
    ts = tgt[order]
    cnt = np.bincount(tgt, minlength=n_rows).astype(np.int32)
    assert cnt.max(initial=0) <= 4, "a target has more than 4 in-edges"
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    pos = np.arange(E) - start[ts]
    nbr = np.full((n_rows, 4), -1, dtype=np.int32)
    eid = np.full((n_rows, 4), -1, dtype=np.int32)
    nbr[ts, pos] = src[order]
    eid[ts, pos] = order
    return nbr, eid, cnt


def morton_codes(pos: np.ndarray, bits: int = 10) -> np.ndarray:
    """30-bit (default) Morton code of each position, quantised over the bounding box
    in float32 arithmetic (the exact formula the device builder uses)."""
    p = pos.astype(np.float32)
    lo = p.min(axis=0)
    hi = p.max(axis=0)
    ext = np.maximum(hi - lo, np.float32(1e-30)).astype(np.float32)
    scale = np.float32((1 << bits) - 1)
    q = ((p - lo) / ext * scale).astype(np.float32)
    q = np.clip(q, 0, scale).astype(np.uint32)
    code = np.zeros(p.shape[0], dtype=np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a].astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + (2 - a))
    return code


def morton_perm(pos: np.ndarray, bits: int = 10) -> np.ndarray:
    """``perm[new] = old``: stable sort of cells by Morton code."""
    return np.argsort(morton_codes(pos, bits), kind="stable").astype(np.int32)


def invert_perm(perm: np.ndarray) -> np.ndarray:
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.shape[0], dtype=perm.dtype)
    return inv


def apply_perm_ell(nbr: np.ndarray, perm: np.ndarray) -> np.ndarray:
    """Renumber an ELL table: row ``new`` is old row ``perm[new]`` with ids mapped by inv."""
    inv = invert_perm(perm)
    out = nbr[perm]
    valid = out >= 0
    out2 = out.copy()
    out2[valid] = inv[out[valid]]
    return out2.astype(np.int32)


def relayout_edges(edge_attr: np.ndarray, nbr: np.ndarray, rslot: np.ndarray, perm: np.ndarray | None = None):
    """Edge attributes in *incoming* and *own-slot* order for (optionally permuted) cells.

    ``ea_in[t,k]  = edge_attr[4*nbr[t,k] + rslot[t,k]]`` (edge nbr->t, used by the forward)
    ``ea_own[s,k] = edge_attr[4*s + k]``                 (edge s->nbr, used by the backward)
    With ``perm`` the rows are those of old cell ``perm[new]`` (old-id ``nbr``/``rslot``).
    """
    N = nbr.shape[0]
    rows_in = 4 * nbr.astype(np.int64) + rslot.astype(np.int64)
    rows_own = 4 * np.arange(N, dtype=np.int64)[:, None] + np.arange(4)[None, :]
    if perm is not None:
        rows_in = rows_in[perm]
        rows_own = rows_own[perm]
    return edge_attr[rows_in], edge_attr[rows_own]


# --------------------------------------------------------------------------- partition / halo


def partition_ranges(n_cells: int, n_parts: int) -> np.ndarray:
    """Contiguous split of (Morton-ordered) cells: ``bounds[p]..bounds[p+1]``."""
    return (np.arange(n_parts + 1, dtype=np.int64) * n_cells // n_parts).astype(np.int64)


def halo_maps(nbr: np.ndarray, bounds: np.ndarray, part: int):
    """Halo maps of one rank for a 1-ring exchange.

    ``nbr`` is the global ELL table in partition order.  Returns a dict with
      ``local_nbr``  int32[n_own,4]   ids into ``[owned | halo]`` rows,
      ``halo_gid``   int64[n_halo]    global ids of halo rows, ascending,
      ``recv_counts``int64[P]         halo rows owned by each peer (contiguous, peer order),
      ``send_idx``   list of int32[]  per peer: local owned rows that peer needs, ascending.
    """
    P = bounds.shape[0] - 1
    lo, hi = int(bounds[part]), int(bounds[part + 1])
    own = nbr[lo:hi]
    remote = (own < lo) | (own >= hi)
    halo_gid = np.unique(own[remote].astype(np.int64))
    local = np.where(remote, 0, own - lo).astype(np.int64)
    if halo_gid.size:
        local[remote] = (hi - lo) + np.searchsorted(halo_gid, own[remote].astype(np.int64))
    owner = np.searchsorted(bounds, halo_gid, side="right") - 1
    recv_counts = np.bincount(owner, minlength=P).astype(np.int64)
    send_idx = []
    for q in range(P):
        if q == part:
            send_idx.append(np.zeros(0, dtype=np.int32))
            continue
        qlo, qhi = int(bounds[q]), int(bounds[q + 1])
        theirs = nbr[qlo:qhi]
        need = np.unique(theirs[(theirs >= lo) & (theirs < hi)].astype(np.int64))
        send_idx.append((need - lo).astype(np.int32))
    return dict(local_nbr=local.astype(np.int32), halo_gid=halo_gid, recv_counts=recv_counts,
                send_idx=send_idx)


# --------------------------------------------------------------------------- labels / facets


def labels_from_logits(logits: np.ndarray) -> np.ndarray:
    """``processing/generate_mesh.py:75``: argmax of log_softmax, ties -> 0 (inside)."""
    z = logits.astype(np.float32)
    return (z[:, 1] > z[:, 0]).astype(np.int64)


def interface_facets(labels_finite: np.ndarray, nfacets: np.ndarray) -> np.ndarray:
    """``processing/generate_mesh.py:94-105``: facets whose two cells differ, the
    infinite cell (-1) forced outside (label 1)."""
    lab = np.append(labels_finite, 1)
    e = np.where(nfacets < 0, labels_finite.shape[0], nfacets)
    return np.nonzero(lab[e[:, 0]] != lab[e[:, 1]])[0]
