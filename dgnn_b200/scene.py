"""Sharded scenes: what ONE rank holds of a cell graph that is partitioned over the GPUs of a box
(BASELINE configs[3],[4]).

A ``LocalScene`` is a rank's contiguous range ``[lo, hi)`` of the partition (Morton) order with, for its owned cells
only, the ELL-4 rows in GLOBAL partition-order ids, the node features, the edge attributes in incoming order
(``ea_in[t,k]`` = attributes of the edge ``nbr[t,k] -> t``, forward) and own-slot order (``ea_own[s,k]`` = attributes of
``s -> nbr[s,k]``, backward), and the supervision.  No rank ever materialises the whole scene:
``dgnn_b200.partition.build_halo_maps_sharded`` finds the halo from these rows and the peers.

``lattice_scene`` generates such a shard analytically on the device for the largest benchmark configs (SURVEY.md
section 7, "hard parts": scipy Delaunay of 10 M points is out of reach, an analytic 4-regular lattice is allowed):
the diamond-cubic lattice, every site with 4 neighbours like the cells of a tetrahedralisation, periodic, power-of-two
extents, cells numbered along the Morton curve of the lattice cubes.  Features are a counter-based hash of the GLOBAL
cell / edge id, so a scene is the same whatever the number of ranks - which is what lets a partitioned run be compared
with a single-GPU run.  ``scene_from_global`` cuts a shard out of an ordinary (host) graph in the reference layout.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch


@dataclass
class LocalScene:
    n_global: int
    lo: int
    hi: int
    nbr_gid: torch.Tensor                 # int64[n_own,4] global ids (partition order), -1 = none
    x: torch.Tensor                       # float32[n_own, F0] model input columns (the weight column already dropped)
    ea_in: Optional[torch.Tensor]         # float32[n_own,4,Fe]
    ea_own: Optional[torch.Tensor] = None  # float32[n_own,4,Fe] (training)
    y: Optional[torch.Tensor] = None      # float32[n_own,2]
    w: Optional[torch.Tensor] = None      # float32[n_own] raw loss weight (column 0 of the reference's x)
    caller_ids: Optional[torch.Tensor] = None   # int64[n_own] ids of the owned cells in the caller's numbering

    @property
    def n_own(self) -> int:
        return self.hi - self.lo


# --------------------------------------------------------------------------- counter-based features


def _s64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


_M1, _M2, _M3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def hash_uniform(idx: torch.Tensor, salt: int) -> torch.Tensor:
    """float32 in [0, 1) from int64 counters (splitmix64 finaliser; identical on CPU and GPU)."""
    z = idx * _M1 + _s64(salt * 0x632BE59BD9B4E019 + 0x1234567)
    z = (z ^ ((z >> 30) & 0x3FFFFFFFF)) * _M2
    z = (z ^ ((z >> 27) & 0x1FFFFFFFFF)) * _M3
    z = z ^ ((z >> 31) & 0x1FFFFFFFF)
    return (z & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)


def hash_normalish(idx: torch.Tensor, salt: int) -> torch.Tensor:
    """Zero-mean, unit-variance values (sum of three uniforms) from int64 counters."""
    u = hash_uniform(idx, salt) + hash_uniform(idx, salt + 7919) + hash_uniform(idx, salt + 15838)
    return (u - 1.5) * 2.0


def cell_features(gid: torch.Tensor, f0: int) -> torch.Tensor:
    """float32[len(gid), f0] node features of the cells ``gid``."""
    cols = torch.arange(f0, device=gid.device, dtype=torch.int64)
    return hash_normalish(gid[:, None] * f0 + cols[None, :], 1)


def edge_features(eid: torch.Tensor, fe: int) -> torch.Tensor:
    """float32[..., fe] attributes of the directed edges ``eid`` (= 4 * owner + slot); -1 -> zeros."""
    cols = torch.arange(fe, device=eid.device, dtype=torch.int64)
    v = hash_normalish(eid.clamp(min=0)[..., None] * fe + cols, 2)
    return torch.where((eid >= 0)[..., None], v, torch.zeros_like(v))


def cell_targets(gid: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(y float32[n,2] = (u, 1-u), w float32[n] raw volume-like weight)."""
    u = hash_uniform(gid, 3)
    return torch.stack([u, 1.0 - u], dim=1), hash_uniform(gid, 4) * 1e-3


# --------------------------------------------------------------------------- diamond lattice along the Morton curve


def _bits(n: int) -> int:
    b = n.bit_length() - 1
    if n <= 0 or (1 << b) != n:
        raise ValueError("lattice extents must be powers of two, got %d" % n)
    return b


def lattice_encode(x, y, z, s, dims):
    """Partition-order id of site (x, y, z, sublattice s): bit-interleaved cube code * 2 + s."""
    bx, by, bz = (_bits(d) for d in dims)
    code = torch.zeros_like(x)
    pos = 0
    for i in range(max(bx, by, bz)):
        for v, b in ((x, bx), (y, by), (z, bz)):
            if i < b:
                code = code | (((v >> i) & 1) << pos)
                pos += 1
    return code * 2 + s


def lattice_decode(gid, dims):
    bx, by, bz = (_bits(d) for d in dims)
    s = gid & 1
    code = gid >> 1
    x = torch.zeros_like(gid); y = torch.zeros_like(gid); z = torch.zeros_like(gid)
    pos = 0
    for i in range(max(bx, by, bz)):
        if i < bx:
            x = x | (((code >> pos) & 1) << i); pos += 1
        if i < by:
            y = y | (((code >> pos) & 1) << i); pos += 1
        if i < bz:
            z = z | (((code >> pos) & 1) << i); pos += 1
    return x, y, z, s


def lattice_neighbours(gid: torch.Tensor, dims) -> torch.Tensor:
    """int64[len(gid),4]: the four bonded sites of every site (periodic diamond-cubic lattice, the bonding of
    ``synthetic.lattice_graph``; slot k of either end of a bond is the same direction, so the reverse slot of k is k)."""
    nx, ny, nz = dims
    x, y, z, s = lattice_decode(gid, dims)
    sgn = 2 * s - 1                                   # sublattice 0 bonds towards -1, sublattice 1 towards +1
    o = 1 - s
    out = [lattice_encode(x, y, z, o, dims),
           lattice_encode((x + sgn) & (nx - 1), y, z, o, dims),
           lattice_encode(x, (y + sgn) & (ny - 1), z, o, dims),
           lattice_encode(x, y, (z + sgn) & (nz - 1), o, dims)]
    return torch.stack(out, dim=1)


def lattice_scene(dims, rank: int, world: int, device, f0: int = 28, fe: int = 20, need_backward: bool = False,
                  chunk: int = 1 << 20) -> LocalScene:
    """This rank's shard of the ``2 * nx * ny * nz``-cell lattice scene, generated on ``device`` in chunks."""
    from .partition import partition_bounds
    n = 2 * dims[0] * dims[1] * dims[2]
    b = partition_bounds(n, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    n_own = hi - lo
    dev = torch.device(device)
    nbr = torch.empty((n_own, 4), dtype=torch.int64, device=dev)
    x = torch.empty((n_own, f0), dtype=torch.float32, device=dev)
    ea_in = torch.empty((n_own, 4, fe), dtype=torch.float32, device=dev) if fe else None
    ea_own = torch.empty((n_own, 4, fe), dtype=torch.float32, device=dev) if (fe and need_backward) else None
    y = torch.empty((n_own, 2), dtype=torch.float32, device=dev)
    w = torch.empty(n_own, dtype=torch.float32, device=dev)
    slot = torch.arange(4, device=dev, dtype=torch.int64)[None, :]
    for c0 in range(0, n_own, chunk):
        c1 = min(n_own, c0 + chunk)
        gid = torch.arange(lo + c0, lo + c1, device=dev, dtype=torch.int64)
        nb = lattice_neighbours(gid, dims)
        nbr[c0:c1] = nb
        x[c0:c1] = cell_features(gid, f0)
        if fe:
            ea_in[c0:c1] = edge_features(nb * 4 + slot, fe)            # edge nbr[t,k] -> t is the neighbour's slot k
            if ea_own is not None:
                ea_own[c0:c1] = edge_features(gid[:, None] * 4 + slot, fe)
        y[c0:c1], w[c0:c1] = cell_targets(gid)
    return LocalScene(n, lo, hi, nbr, x, ea_in, ea_own, y, w,
                      caller_ids=torch.arange(lo, hi, device=dev, dtype=torch.int64))


def lattice_global(dims, f0: int = 28, fe: int = 20):
    """The whole lattice scene as an ordinary host graph in the reference layout (cells numbered in partition order):
    ``dict(x[N,1+f0], edge_attr[4N,fe], edge_index[2,4N], y[N,2], pos[N,3])`` - for checking a partitioned run against
    the single-GPU path on small extents."""
    n = 2 * dims[0] * dims[1] * dims[2]
    gid = torch.arange(n, dtype=torch.int64)
    nb = lattice_neighbours(gid, dims)
    y, w = cell_targets(gid)
    x = torch.cat([w[:, None], cell_features(gid, f0)], dim=1)
    ea = edge_features(torch.arange(4 * n, dtype=torch.int64), fe) if fe else None
    ei = torch.stack([gid.repeat_interleave(4), nb.reshape(-1)])
    cx, cy, cz, s = lattice_decode(gid, dims)
    pos = torch.stack([cx, cy, cz], dim=1).to(torch.float32) + 0.25 + 0.5 * s[:, None].to(torch.float32)
    return dict(x=x, edge_attr=ea, edge_index=ei, y=y, pos=pos)


# --------------------------------------------------------------------------- a shard of an ordinary graph


def scene_from_global(data_all, rank: int, world: int, device, cell_type=True, edge_type=False, use_edges=True,
                      need_backward: bool = False) -> LocalScene:
    """Cut this rank's shard out of a whole graph in the reference layout held on the HOST (``x[N,1+F0]``,
    ``edge_attr[4N,Fe]``, ``edge_index[2,4N]`` with row ``4i+k`` = facet ``k`` of cell ``i``, optional ``pos``).

    Only the ordering works on whole-graph arrays (positions and the adjacency, 28 B per cell); features and edge
    attributes are gathered for the owned cells alone, so the device never holds more than its shard."""
    from .graph import locality_order
    from .partition import partition_bounds
    dev = torch.device(device)
    n = data_all.x.shape[0]
    ei = data_all.edge_index
    if ei.shape[1] != 4 * n:
        raise ValueError("scene_from_global expects the reference file layout (4 facet rows per cell)")
    adj = ei[1].reshape(n, 4).to(torch.int64)                      # facet neighbours in caller numbering (host)
    pos = getattr(data_all, "pos", None)
    perm = locality_order(ei.cpu() if pos is None else None, n, pos, dev)   # int32[n] new -> old, on the device
    perm = perm.long().cpu() if perm is not None else torch.arange(n)
    inv = torch.empty(n, dtype=torch.int64)
    inv[perm] = torch.arange(n)
    b = partition_bounds(n, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    own_old = perm[lo:hi]                                          # caller ids of the owned cells
    nb_old = adj[own_old]                                          # [n_own,4] caller ids of their neighbours
    nbr_gid = inv[nb_old]
    cols = slice(1, None) if cell_type else slice(None)
    x = data_all.x[own_old][:, cols].to(dev, dtype=torch.float32)
    ea_in = ea_own = None
    if use_edges and data_all.edge_attr is not None:
        ea = data_all.edge_attr[:, 1:] if edge_type else data_all.edge_attr
        # reverse slot: the facet j of the neighbour that points back at the owned cell
        back = adj[nb_old.reshape(-1)].reshape(-1, 4, 4) == own_old[:, None, None]
        if not bool(back.any(dim=2).all()):
            raise ValueError("adjacency is not symmetric")
        rslot = back.to(torch.int64).argmax(dim=2)
        ea_in = ea[(nb_old * 4 + rslot).reshape(-1)].reshape(hi - lo, 4, -1).to(dev, dtype=torch.float32)
        if need_backward:
            ea_own = ea[(own_old[:, None] * 4 + torch.arange(4)[None, :]).reshape(-1)].reshape(hi - lo, 4, -1).to(
                dev, dtype=torch.float32)
    y = data_all.y[own_old].to(dev, dtype=torch.float32) if getattr(data_all, "y", None) is not None else None
    w = data_all.x[own_old][:, 0].to(dev, dtype=torch.float32) if cell_type else None
    return LocalScene(n, lo, hi, nbr_gid.to(dev), x.contiguous(), ea_in, ea_own, y, w, caller_ids=own_old.to(dev))
