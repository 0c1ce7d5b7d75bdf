"""Graph loader re-layout: the reference's ``edge_index`` / ``edge_attr``
(``processing/data.py:434-439,512-519``; ``run.py:211-213``) to the device-resident ELL-4 layout
the fused kernels consume.

``EllGraph`` holds, for one (bipartite) message-passing step with ``n_src`` source rows and
``n_tgt <= n_src`` target rows (targets are the first ``n_tgt`` sources, PyG convention,
``surfaceNetStaticEdgeFilters.py:217``):

* ``nbr   int32[n_tgt,4]``     source row of the k-th in-edge (-1 = none)
* ``ea_in float32[n_tgt,4,Fe]`` that edge's attributes (incoming order, forward)
* ``onbr  int32[n_src,4]``     target row of the k-th out-edge (-1 = none)
* ``ea_own float32[n_src,4,Fe]`` that edge's attributes (own-slot order, backward)

For a whole Delaunay graph in the reference file layout (row ``4i+k`` = facet ``k`` of cell ``i``)
the table is the adjacency itself, ``onbr == nbr`` by symmetry, and the incoming edge row is found
through the reverse-facet slot; cells are renumbered by a locality order (Morton code of the cell
centroid when positions are known, reverse Cuthill-McKee of the adjacency otherwise).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def pad_cols(t: torch.Tensor, width: int) -> torch.Tensor:
    """Zero-pad the last dim of a 2-D float tensor to ``width`` columns (contiguous)."""
    if t.shape[1] == width:
        return t.contiguous()
    out = torch.zeros((t.shape[0], width), dtype=t.dtype, device=t.device)
    out[:, :t.shape[1]] = t
    return out


@dataclass
class EllGraph:
    n_src: int
    n_tgt: int
    fe: int                                  # padded edge feature width (0 = no edge features)
    nbr: torch.Tensor
    ea_in: Optional[torch.Tensor]
    onbr: Optional[torch.Tensor] = None
    ea_own: Optional[torch.Tensor] = None
    perm: Optional[torch.Tensor] = None      # int32[n]: new -> old  (None = caller order)
    inv: Optional[torch.Tensor] = None       # int32[n]: old -> new
    _eid_in: Optional[torch.Tensor] = None   # int32[n_tgt,4]: local edge id of every (target, slot), -1 = none
    _eid_out: Optional[torch.Tensor] = None  # int32[n_src,4]: local edge id of every (source, out-slot), -1 = none
    ea_edges: Optional[torch.Tensor] = None  # float32[E, fe]: edge attributes in edge-list order (edge_convs == 2)
    _orow: Optional[torch.Tensor] = None     # int32[n_src,4]: row 4t+k (incoming order) of every out-edge, -1 = none
    self_loops: bool = False                 # every target is also its own in-neighbour (edge_convs == 0, run.py:70-71)

    def orow(self) -> torch.Tensor:
        """Row of the incoming-order edge matrix [n_tgt*4, .] that holds the out-edge (s, j): inverts _eid_in."""
        if self._orow is None:
            dev = self._eid_in.device
            flat = self._eid_in.reshape(-1)
            ok = flat >= 0
            n_edges = int(flat.max().item()) + 1 if flat.numel() else 0
            row_of_edge = torch.full((max(n_edges, 1),), -1, dtype=torch.int32, device=dev)
            row_of_edge[flat[ok].long()] = torch.arange(flat.numel(), dtype=torch.int32, device=dev)[ok]
            eo = self._eid_out
            self._orow = torch.where(eo >= 0, row_of_edge[eo.clamp(min=0).long()], torch.full_like(eo, -1)).contiguous()
        return self._orow

    def permute_rows(self, x: torch.Tensor) -> torch.Tensor:
        """Rows of ``x`` (caller order, width multiple of 4) in internal order."""
        if self.perm is None:
            return x.contiguous()
        out = torch.empty_like(x)
        call("dgnn_gather_rows", ptr(x), ptr(self.perm), x.shape[0], x.shape[1], ptr(out), _stream())
        return out

    def unpermute_rows(self, y: torch.Tensor) -> torch.Tensor:
        """Rows of ``y`` (internal order) back in caller order."""
        if self.perm is None:
            return y
        if y.shape[1] % 4 == 0:
            out = torch.empty_like(y)
            call("dgnn_gather_rows", ptr(y), ptr(self.inv), y.shape[0], y.shape[1], ptr(out), _stream())
            return out
        return y.index_select(0, self.inv.long())


def is_reference_layout(edge_index: torch.Tensor, n: int) -> bool:
    """True if ``edge_index`` is the reference file layout: E == 4n and row 4i+k is owned by i."""
    if edge_index.shape[1] != 4 * n or n == 0:
        return False
    own = edge_index[0].view(n, 4)
    return bool((own == torch.arange(n, device=own.device, dtype=own.dtype)[:, None]).all())


def locality_order(edge_index_cpu: Optional[torch.Tensor], n: int, pos: Optional[torch.Tensor], device,
                   bits: int = 10) -> Optional[torch.Tensor]:
    """``perm[new] = old`` as int32 on ``device``: Morton order of ``pos`` if given, else reverse
    Cuthill-McKee of the adjacency (host, scipy), else None."""
    if pos is not None:
        p = pos.to(device=device, dtype=torch.float32).contiguous()
        lo = p.min(dim=0).values.cpu().numpy().astype(np.float32)
        hi = p.max(dim=0).values.cpu().numpy().astype(np.float32)
        codes = torch.empty(n, dtype=torch.int64, device=device)
        call("dgnn_morton_codes", ptr(p), n, lo.ctypes.data, hi.ctypes.data, bits, ptr(codes), _stream())
        # stable sort of the codes (graph preparation, not the hot path)
        return torch.sort(codes, stable=True).indices.to(torch.int32)
    if edge_index_cpu is not None:
        import scipy.sparse as sp
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        ei = edge_index_cpu.numpy()
        a = sp.csr_matrix((np.ones(ei.shape[1], dtype=np.int8), (ei[1], ei[0])), shape=(n, n))
        perm = reverse_cuthill_mckee(a, symmetric_mode=True)
        return torch.from_numpy(np.ascontiguousarray(perm.astype(np.int32))).to(device)
    return None


def strip_self_loops(edge_index: torch.Tensor, n_tgt: int, edge_attr) -> Tuple[torch.Tensor, bool]:
    """``add_self_loops`` (``run.py:70-71,215-216``, only when ``model.edge_convs == 0``) appends one edge ``t -> t`` per
    node.  The ELL-4 table keeps the four facet neighbours; the self edge becomes a flag of the layout (the kernels add
    ``h(t)`` to the sum and 1 to the count).  Returns the edge list without self edges and whether there were any; a
    graph where only SOME targets carry a self edge, or one with edge attributes, is rejected."""
    if edge_index.shape[1] < n_tgt:                 # add_self_loops appends one edge per node: fewer edges, no self loops
        return edge_index, False
    loops = edge_index[0] == edge_index[1]
    n_loops = int(loops.sum().item())
    if n_loops == 0:
        return edge_index, False
    if edge_attr is not None:
        raise _lib.DgnnError("self loops are only defined without edge attributes (model.edge_convs == 0, run.py:70)")
    tg = edge_index[1][loops]
    if n_loops != n_tgt or int(torch.unique(tg).numel()) != n_tgt or int(tg.max().item()) >= n_tgt:
        raise _lib.DgnnError("self loops must cover every target exactly once (add_self_loops)")
    return edge_index[:, ~loops].contiguous(), True


def build_full_graph(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor], n: int, device,
                     pos: Optional[torch.Tensor] = None, order: str = "auto",
                     need_backward: bool = True, e_id: Optional[torch.Tensor] = None) -> EllGraph:
    """Whole-graph layout (inference, or training on whole graphs).

    ``order``: "auto" (Morton if ``pos`` else RCM), "morton", "rcm", "none".
    ``e_id`` (int64[4n], optional): edge r of the graph carries row ``e_id[r]`` of ``edge_attr``; the selection
    is fused into the re-layout kernel instead of materialising ``edge_attr[e_id]``.
    """
    dev = torch.device(device)
    _lib.check_device(dev.index or 0)
    # (4 n edges = the facet list of the reference's files, where a cell never neighbours itself: no scan, no sync)
    loops = False
    if edge_index.shape[1] != 4 * n:
        edge_index, loops = strip_self_loops(edge_index, n, edge_attr)
    if loops:
        g = build_full_graph(edge_index, None, n, device, pos, order, need_backward, None)
        g.self_loops = True
        return g
    fe = 0 if edge_attr is None else pad4(edge_attr.shape[1])
    if e_id is not None and edge_attr is not None and (edge_attr.shape[1] % 4 or not edge_attr.is_contiguous()
                                                       or edge_attr.dtype != torch.float32):
        edge_attr, e_id = edge_attr.to(dev)[e_id.to(dev)], None      # odd widths: select first, pad below
    if not is_reference_layout(edge_index, n):
        g = build_from_edges(edge_index.to(dev), e_id, edge_attr, n, n, dev, need_backward)
        return g
    st = _stream()
    adj = edge_index.t().to(torch.int32).contiguous().to(dev)  # [4n,2] == adjacencies.npz layout
    nbr0 = torch.empty((n, 4), dtype=torch.int32, device=dev)
    rslot = torch.empty((n, 4), dtype=torch.uint8, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    call("dgnn_ell_from_adjacency", ptr(adj), n, ptr(nbr0), ptr(rslot), ptr(err), st)
    code = int(err.item())
    if code == 2:  # not symmetric: the generic edge-list builder handles it
        return build_from_edges(edge_index.to(dev), e_id, edge_attr, n, n, dev, need_backward)
    if code != 0:
        raise _lib.DgnnError("dgnn_ell_from_adjacency: malformed adjacency (code %d)" % code)
    perm = None
    if order != "none":
        use_pos = pos if order in ("auto", "morton") else None
        if order == "morton" and pos is None:
            raise ValueError("order='morton' needs positions")
        ei_cpu = edge_index.cpu() if (use_pos is None and order in ("auto", "rcm")) else None
        perm = locality_order(ei_cpu, n, use_pos, dev)
    if perm is not None:
        inv = torch.empty_like(perm)
        inv[perm.long()] = torch.arange(n, dtype=torch.int32, device=dev)
        nbr = torch.empty_like(nbr0)
        call("dgnn_perm_apply_ell", ptr(nbr0), ptr(perm), ptr(inv), n, ptr(nbr), st)
    else:
        inv = None
        nbr = nbr0
    ea_in = ea_own = None
    if edge_attr is not None:
        ea = pad_cols(edge_attr.to(dev, dtype=torch.float32), fe)
        ea_in = torch.empty((n, 4, fe), dtype=torch.float32, device=dev)
        own_copy = need_backward and (perm is not None or e_id is not None)
        if need_backward:
            ea_own = torch.empty((n, 4, fe), dtype=torch.float32, device=dev) if own_copy else ea.view(n, 4, fe)
        if e_id is not None:
            eid = e_id.to(dev, dtype=torch.int64).contiguous()
            call("dgnn_edge_relayout_idx", ptr(ea), ptr(eid), ptr(nbr0), ptr(rslot), ptr(perm), n, fe, ptr(ea_in),
                 ptr(ea_own) if own_copy else None, st)
        else:
            call("dgnn_edge_relayout", ptr(ea), ptr(nbr0), ptr(rslot), ptr(perm), n, fe, ptr(ea_in),
                 ptr(ea_own) if own_copy else None, st)
    return EllGraph(n_src=n, n_tgt=n, fe=fe, nbr=nbr, ea_in=ea_in, onbr=nbr if need_backward else None,
                    ea_own=ea_own, perm=perm, inv=inv)


def build_from_edges(edge_index: torch.Tensor, e_id: Optional[torch.Tensor], edge_attr: Optional[torch.Tensor],
                     n_src: int, n_tgt: int, device, need_backward: bool = True) -> EllGraph:
    """Generic (bipartite / sampled) layout from an edge list in local ids
    (``data.batch_adjs[i] = (edge_index, e_id, size)``, ``surfaceNetStaticEdgeFilters.py:215-217``).
    ``edge_attr`` rows are selected by ``e_id`` (all edges in order when ``e_id`` is None)."""
    dev = torch.device(device)
    _lib.check_device(dev.index or 0)
    edge_index, loops = strip_self_loops(edge_index, n_tgt, edge_attr)
    if loops:
        g = build_from_edges(edge_index, None, None, n_src, n_tgt, device, need_backward)
        g.self_loops = True
        return g
    st = _stream()
    ei = edge_index.to(dev, dtype=torch.int64).contiguous()
    E = ei.shape[1]
    fe = 0 if edge_attr is None else pad4(edge_attr.shape[1])
    err = torch.zeros(1, dtype=torch.int32, device=dev)

    def one_side(src, tgt, rows, others):
        nb = torch.empty((rows, 4), dtype=torch.int32, device=dev)
        eid = torch.empty((rows, 4), dtype=torch.int32, device=dev)
        cnt = torch.zeros(rows, dtype=torch.int32, device=dev)
        call("dgnn_ell_build", ptr(src), ptr(tgt), E, rows, others, 0, ptr(nb), ptr(eid), ptr(cnt), ptr(err), st)
        return nb, eid

    nbr, eid_in = one_side(ei[0], ei[1], n_tgt, n_src)
    onbr = eid_out = None
    if need_backward:
        onbr, eid_out = one_side(ei[1], ei[0], n_src, n_tgt)
    code = int(err.item())
    if code == 3:
        raise _lib.DgnnError("a cell has more than 4 facet neighbours: not a Delaunay cell graph "
                             "(self-loops / cliques are not supported by the ELL-4 layout)")
    if code != 0:
        raise _lib.DgnnError("dgnn_ell_build: edge endpoint out of range (code %d)" % code)
    ea_in = ea_own = ea = None
    if edge_attr is not None:
        if e_id is None:
            rows = edge_attr
        elif e_id.numel() * 2 >= edge_attr.shape[0]:
            rows = edge_attr.to(dev, non_blocking=True)[e_id.to(dev)]
        else:
            rows = edge_attr[e_id.to(edge_attr.device)]
        ea = pad_cols(rows.to(dev, dtype=torch.float32), fe)
        ea_in = torch.empty((n_tgt, 4, fe), dtype=torch.float32, device=dev)
        call("dgnn_gather_rows", ptr(ea), ptr(eid_in), n_tgt * 4, fe, ptr(ea_in), st)
        if need_backward:
            ea_own = torch.empty((n_src, 4, fe), dtype=torch.float32, device=dev)
            call("dgnn_gather_rows", ptr(ea), ptr(eid_out), n_src * 4, fe, ptr(ea_own), st)
    return EllGraph(n_src=n_src, n_tgt=n_tgt, fe=fe, nbr=nbr, ea_in=ea_in, onbr=onbr, ea_own=ea_own, _eid_in=eid_in,
                    _eid_out=eid_out, ea_edges=ea)
