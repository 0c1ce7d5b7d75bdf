"""Spatial partition of a cell graph across the GPUs of one box with a per-layer halo exchange
(north_star subsystem 4).

Cells are ordered by the Morton code of their centroid and split into ``world_size`` contiguous
ranges.  A rank owns one range and keeps, behind its owned rows, a *halo*: the remote cells adjacent
to owned cells, grouped by owning rank (ascending global id).  Before every layer after the first,
the boundary rows each peer needs are packed (``dgnn_gather_rows``), exchanged with ONE
``all_to_all_single`` (NCCL over NVLink; gloo in the CPU tests) and land directly in the halo rows of
the activation matrix, so the local ELL table indexes ``[owned | halo]`` without further copies.

The maps are integer tables; ``tests/test_partition_cpu.py`` checks them bit-exactly against the NumPy
oracle and runs the exchange with world_size 2 on gloo.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class HaloMaps:
    rank: int
    world: int
    lo: int
    hi: int
    n_own: int
    n_halo: int
    local_nbr: torch.Tensor          # int32[n_own,4] ids into [owned | halo]
    halo_gid: torch.Tensor           # int64[n_halo] global ids (partition order), ascending = grouped by owner
    recv_counts: List[int]           # halo rows owned by each peer
    send_idx: torch.Tensor           # int32[sum(send_counts)] local owned rows, grouped by destination peer
    send_counts: List[int]
    n_boundary: int = 0              # boundary-first order: owned rows [0, n_boundary) are the ones some peer needs
    order: Optional[torch.Tensor] = None   # int64[n_own]: new local row -> row before the boundary-first reorder


def partition_bounds(n_cells: int, world: int) -> torch.Tensor:
    return (torch.arange(world + 1, dtype=torch.int64) * n_cells) // world


def build_halo_maps(nbr: torch.Tensor, bounds: torch.Tensor, rank: int) -> HaloMaps:
    """``nbr`` int32[N,4]: the global ELL table in partition (Morton) order, on any device.
    Every rank holds the whole table (replicated build, sharded compute)."""
    world = bounds.numel() - 1
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    dev = nbr.device
    own = nbr[lo:hi].to(torch.int64)
    remote = (own < lo) | (own >= hi)
    halo_gid = torch.unique(own[remote])                       # sorted ascending
    local = own - lo
    if halo_gid.numel():
        local = torch.where(remote, (hi - lo) + torch.searchsorted(halo_gid, own), local)
    b = bounds.to(dev)
    owner = torch.searchsorted(b, halo_gid, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=world).tolist() if halo_gid.numel() else [0] * world
    send, send_counts = [], []
    for q in range(world):
        if q == rank:
            send_counts.append(0)
            continue
        theirs = nbr[int(bounds[q]):int(bounds[q + 1])].to(torch.int64)
        need = torch.unique(theirs[(theirs >= lo) & (theirs < hi)]) - lo
        send.append(need.to(torch.int32))
        send_counts.append(int(need.numel()))
    send_idx = torch.cat(send) if send else torch.zeros(0, dtype=torch.int32, device=dev)
    return HaloMaps(rank, world, lo, hi, hi - lo, int(halo_gid.numel()), local.to(torch.int32).contiguous(), halo_gid,
                    [int(c) for c in recv_counts], send_idx.contiguous(), send_counts)


def build_halo_maps_sharded(own_nbr: torch.Tensor, lo: int, hi: int, bounds: torch.Tensor, rank: int,
                            group=None) -> HaloMaps:
    """Same maps as ``build_halo_maps`` from this rank's rows of the table only (``own_nbr`` int64/int32[n_own,4], global
    ids in partition order, -1 = no neighbour): no rank ever holds the global table.  The halo ids are found locally;
    which owned rows every peer needs is learnt from the peers themselves with two all-to-alls (counts, then ids) -
    O(halo) communication.  Collective over ``group`` (NCCL for CUDA tensors, gloo for CPU tensors)."""
    world = bounds.numel() - 1
    dev = own_nbr.device
    own = own_nbr.to(torch.int64)
    n_own = hi - lo
    valid = own >= 0
    remote = valid & ((own < lo) | (own >= hi))
    halo_gid = torch.unique(own[remote])                       # sorted ascending => grouped by owner
    local = torch.where(valid, own - lo, own)
    if halo_gid.numel():
        local = torch.where(remote, n_own + torch.searchsorted(halo_gid, own.clamp(min=0)), local)
    b = bounds.to(dev)
    owner = torch.searchsorted(b, halo_gid, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=world) if halo_gid.numel() else torch.zeros(world, dtype=torch.int64, device=dev)
    if world == 1:
        return HaloMaps(rank, 1, lo, hi, n_own, 0, local.to(torch.int32).contiguous(), halo_gid, [0],
                        torch.zeros(0, dtype=torch.int32, device=dev), [0])
    send_counts = torch.empty_like(recv_counts)
    dist.all_to_all_single(send_counts, recv_counts.contiguous(), group=group)   # how many of MY rows each peer asks for
    rc, sc = [int(v) for v in recv_counts.tolist()], [int(v) for v in send_counts.tolist()]
    wanted = torch.empty(sum(sc), dtype=torch.int64, device=dev)
    dist.all_to_all_single(wanted, halo_gid.contiguous(), output_split_sizes=sc, input_split_sizes=rc, group=group)
    send_idx = (wanted - lo).to(torch.int32).contiguous()
    return HaloMaps(rank, world, lo, hi, n_own, int(halo_gid.numel()), local.to(torch.int32).contiguous(), halo_gid, rc,
                    send_idx, sc)


def boundary_first(m: HaloMaps) -> HaloMaps:
    """Renumber the owned rows as [boundary | interior] (boundary = rows some peer needs), so that a layer can compute
    the boundary rows first, start the exchange on the communication stream and compute the interior while the halo
    rows fly (SURVEY 8e "Overlap").  Within each class the previous order (Morton) is kept.  ``order[new] = old``."""
    dev = m.local_nbr.device
    is_b = torch.zeros(m.n_own, dtype=torch.bool, device=dev)
    if m.send_idx.numel():
        is_b[m.send_idx.long()] = True
    order = torch.cat([is_b.nonzero().view(-1), (~is_b).nonzero().view(-1)])
    inv = torch.empty(m.n_own, dtype=torch.int64, device=dev)
    inv[order] = torch.arange(m.n_own, device=dev)
    nb = m.local_nbr.long()[order]
    owned = (nb >= 0) & (nb < m.n_own)
    nb = torch.where(owned, inv[nb.clamp(min=0, max=max(m.n_own - 1, 0))], nb)
    return HaloMaps(m.rank, m.world, m.lo, m.hi, m.n_own, m.n_halo, nb.to(torch.int32).contiguous(), m.halo_gid,
                    m.recv_counts, inv[m.send_idx.long()].to(torch.int32).contiguous(), m.send_counts,
                    n_boundary=int(is_b.sum().item()), order=order)


def exchange_halo(h: torch.Tensor, m: HaloMaps, group=None, send: Optional[torch.Tensor] = None) -> None:
    """Fill the halo rows ``h[n_own:]`` with the owners' rows.  ``h``: float32[n_own + n_halo, F].
    ``send``: optional persistent pack buffer with at least ``len(send_idx) * F`` elements."""
    if m.world == 1:
        return
    f = h.shape[1]
    n_send = int(m.send_idx.numel())
    if h.is_cuda:
        from ._lib import call, ptr
        if send is None or send.numel() < n_send * f:
            send = torch.empty(max(n_send * f, 1), dtype=h.dtype, device=h.device)
        send = send[:n_send * f].view(n_send, f)
        if n_send:
            call("dgnn_gather_rows", ptr(h), ptr(m.send_idx), n_send, f, ptr(send), torch.cuda.current_stream().cuda_stream)
    else:  # host logic of the CPU tests (gloo)
        send = h.index_select(0, m.send_idx.long())
    recv = h[m.n_own:]                                        # halo rows are contiguous and grouped by owner
    dist.all_to_all_single(recv, send, output_split_sizes=m.recv_counts, input_split_sizes=m.send_counts, group=group)


class HaloComm:
    """What ``engine.forward`` / ``engine.backward`` need from a partition: the halo exchange (activations
    forward, d_agg backward - the adjacency is symmetric, so it is the same exchange), a sum over ranks for the
    normalisation statistics, and the number of cells of all ranks.

    With boundary-first maps (``maps.n_boundary > 0``) the exchange can be split from the computation:
    ``start(h)`` - the boundary rows of ``h`` are complete on the compute stream - packs them and runs the all-to-all on
    a communication stream; ``finish()`` makes the compute stream wait for the halo rows.  The pack buffer is persistent
    (sized for the widest layer seen)."""

    def __init__(self, maps: HaloMaps, n_rows: int, group=None):
        self.maps, self.n_rows, self.group = maps, int(n_rows), group
        self._send = None
        self._stream = None
        self._done = None
        #: seconds-resolution bookkeeping for the benchmark: (exchanges, rows sent, bytes sent)
        self.stats = {"exchanges": 0, "bytes_sent": 0}

    @property
    def n_boundary(self) -> int:
        return self.maps.n_boundary if self.maps.world > 1 else 0

    def _buf(self, h: torch.Tensor) -> torch.Tensor:
        need = int(self.maps.send_idx.numel()) * h.shape[1]
        if self._send is None or self._send.numel() < need or self._send.device != h.device:
            self._send = torch.empty(max(need, 1), dtype=h.dtype, device=h.device)
        return self._send

    def exchange(self, h: torch.Tensor) -> None:
        if self.maps.world > 1:
            self.stats["exchanges"] += 1
            self.stats["bytes_sent"] += int(self.maps.send_idx.numel()) * h.shape[1] * h.element_size()
        exchange_halo(h, self.maps, self.group, self._buf(h) if h.is_cuda else None)

    def start(self, h: torch.Tensor) -> None:
        """Asynchronous ``exchange``: everything enqueued so far on the current stream is visible to the exchange."""
        if self.maps.world == 1:
            return
        if not h.is_cuda:
            self.exchange(h)
            return
        cur = torch.cuda.current_stream(h.device)
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=h.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            self.exchange(h)
            self._done = torch.cuda.Event()
            self._done.record(self._stream)

    def finish(self) -> None:
        if self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)
            self._done = None

    def allreduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.maps.world > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def reverse_add(self, d: torch.Tensor) -> None:
        """The mirror of ``exchange`` for gradients that were accumulated on the SOURCE side: the halo rows
        ``d[n_own:]`` hold this rank's contributions to cells owned by peers; they travel back to their owners and are
        added to the owners' rows (one peer at a time, fixed order: deterministic, no atomics)."""
        m = self.maps
        if m.world == 1:
            return
        f = d.shape[1]
        n_back = int(m.send_idx.numel())
        back = torch.empty((n_back, f), dtype=d.dtype, device=d.device)
        dist.all_to_all_single(back, d[m.n_own:].contiguous(), output_split_sizes=m.send_counts,
                               input_split_sizes=m.recv_counts, group=self.group)
        if d.is_cuda:
            from ._lib import call, ptr
            st = torch.cuda.current_stream().cuda_stream
            o = 0
            for q, c in enumerate(m.send_counts):
                if c:
                    call("dgnn_add_rows", ptr(back[o:o + c]), ptr(m.send_idx[o:o + c]), c, f, ptr(d), st)
                o += c
        else:  # host logic of the CPU tests (gloo)
            o = 0
            for c in m.send_counts:
                if c:
                    d.index_add_(0, m.send_idx[o:o + c].long(), back[o:o + c])
                o += c


def _local_plan(net, data_all, world, rank, need_backward):
    """Replicated graph build (Morton order), then this rank's [owned | halo] slice of it."""
    from .graph import EllGraph, build_full_graph, pad4, pad_cols
    dev = net._device()
    cols = slice(1, None) if net.clf.regularization.cell_type else slice(None)
    n = data_all.x.shape[0]
    ea = None
    if net.clf.model.edge_convs:
        ea = data_all.edge_attr[:, 1:] if net.clf.regularization.edge_type else data_all.edge_attr
    pos = getattr(data_all, "pos", None)
    full = build_full_graph(data_all.edge_index.to(torch.long), ea, n, dev, pos=pos, order="auto",
                            need_backward=need_backward)
    bounds = partition_bounds(n, world)
    maps = build_halo_maps(full.nbr, bounds, rank)
    lo, hi = maps.lo, maps.hi
    g = EllGraph(n_src=maps.n_own + maps.n_halo, n_tgt=maps.n_own, fe=full.fe, nbr=maps.local_nbr,
                 ea_in=full.ea_in[lo:hi].contiguous() if full.ea_in is not None else None)
    if need_backward:       # symmetric adjacency: the out-edge table of the owned rows is the in-edge table
        g.onbr = maps.local_nbr
        g.ea_own = full.ea_own[lo:hi].contiguous() if full.ea_own is not None else None
    x = data_all.x[:, cols].to(dev, dtype=torch.float32)
    xp = full.permute_rows(pad_cols(x, pad4(x.shape[1])))
    # layer-0 input for [owned | halo]: features are static, so the halo rows are taken locally
    rows = torch.cat([torch.arange(lo, hi, device=dev), maps.halo_gid.to(dev)])
    x0 = xp.index_select(0, rows).contiguous()
    owned_caller_ids = full.perm[lo:hi].long() if full.perm is not None else torch.arange(lo, hi, device=dev)
    return g, maps, x0, owned_caller_ids, n


def _plan_from_scene(net, scene, rank, world, group, need_backward, overlap=True):
    """[owned | halo] layout of one rank from its ``LocalScene`` alone (sharded build: halo maps negotiated with the peers,
    owned rows renumbered boundary-first, layer-0 features of the halo rows fetched from their owners)."""
    from .graph import EllGraph, pad4, pad_cols
    dev = net._device()
    maps = build_halo_maps_sharded(scene.nbr_gid.to(dev), scene.lo, scene.hi, partition_bounds(scene.n_global, world), rank,
                                   group)
    x, ea_in, ea_own, ids = scene.x, scene.ea_in, scene.ea_own, scene.caller_ids
    if overlap and world > 1:
        maps = boundary_first(maps)
        o = maps.order
        x = x[o]
        ea_in = ea_in[o] if ea_in is not None else None
        ea_own = ea_own[o] if ea_own is not None else None
        ids = ids[o] if ids is not None else None
    n_own, n_src = maps.n_own, maps.n_own + maps.n_halo
    use_edges = bool(net.clf.model.edge_convs) and ea_in is not None
    fe = pad4(ea_in.shape[2]) if use_edges else 0

    def pad_e(e):
        return pad_cols(e.reshape(n_own * 4, -1), fe).view(n_own, 4, fe) if e is not None else None

    g = EllGraph(n_src=n_src, n_tgt=n_own, fe=fe, nbr=maps.local_nbr, ea_in=pad_e(ea_in) if use_edges else None)
    if need_backward:       # symmetric adjacency: the out-edge table of the owned rows is the in-edge table
        g.onbr = maps.local_nbr
        g.ea_own = pad_e(ea_own) if use_edges else None
    comm = HaloComm(maps, scene.n_global, group)
    f0 = pad4(x.shape[1])
    x0 = torch.zeros((n_src, f0), dtype=torch.float32, device=dev)
    x0[:n_own, :x.shape[1]] = x
    comm.exchange(x0)                                  # static input features of the halo rows, once
    return g, maps, x0, ids, comm


class PartitionedTraining:
    """Training of a ``SurfaceNet`` on ONE scene sharded over the ranks of the process group:
    forward halo exchange of the pre-norm activations, normalisation statistics and loss normaliser summed over
    ranks, backward halo exchange of d_agg, and one flat all-reduce of the gradients.

        pt = PartitionedTraining(model)
        ids, logits = pt.forward(data_all)          # logits of the owned cells, autograd-connected
        loss = pt.loss(logits, data_all)            # the GLOBAL loss value (same on every rank)
        optimizer.zero_grad(); loss.backward(); pt.allreduce_gradients(); optimizer.step()
    """

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._plan = None

    def prepare(self, data_all):
        g, maps, x0, ids, n = _local_plan(self.model, data_all, self.world, self.rank, need_backward=True)
        self._plan = (g, maps, x0, ids, HaloComm(maps, n, self.group))
        return self._plan

    def prepare_scene(self, scene, overlap=True):
        """Sharded build: ``scene`` is this rank's ``dgnn_b200.scene.LocalScene`` (needs ``ea_own`` for the backward)."""
        g, maps, x0, ids, comm = _plan_from_scene(self.model, scene, self.rank, self.world, self.group, True, overlap)
        o = maps.order
        self._sup = (scene.y[o] if o is not None else scene.y, scene.w[o] if (o is not None and scene.w is not None) else scene.w)
        self._plan = (g, maps, x0, ids, comm)
        return self._plan

    def forward(self, data_all=None):
        if self._plan is None:
            self.prepare(data_all)
        g, maps, x0, ids, comm = self._plan
        net = self.model
        with torch.cuda.device(x0.device):
            out = net._run([g] * net.num_layers, x0, comm=comm)
        return ids, out

    def loss(self, logits, data_all=None):
        from .runModel import cell_loss
        dist_kw = dict(group=self.group if self.world > 1 else None, distributed=self.world > 1)
        if data_all is None:                              # sharded scene: supervision of the owned cells is local
            y, w = self._sup
            return cell_loss(logits, y, w, self.model.clf, **dist_kw)
        ids = self._plan[3].to(data_all.y.device)
        return cell_loss(logits, data_all.y[ids], data_all.x[ids], self.model.clf, **dist_kw)

    def allreduce_gradients(self):
        ps = [p for p in self.model.parameters() if p.grad is not None]
        if self.world == 1 or not ps:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        dist.all_reduce(flat, group=self.group)
        o = 0
        for p in ps:
            p.grad.copy_(flat[o:o + p.numel()].view_as(p.grad))
            o += p.numel()


class PartitionedUpdatedTraining:
    """BASELINE configs[3]: the Updated-edge-filter model trained on ONE scene sharded over the ranks.  The edge
    state of an edge lives with its target cell, so it never crosses the partition; node activations are exchanged
    forward (halo rows) and the gradients of halo sources travel back to their owners (``HaloComm.reverse_add``).

        pt = PartitionedUpdatedTraining(model)          # dgnn_b200.surfaceNetUpdatedEdgeFilters.SurfaceNet
        ids, out = pt.forward(data_all)                 # data_all: x, edge_index, edge_attr[, pos] of the whole scene
        loss = ...(out, targets[ids]); loss.backward(); pt.allreduce_gradients(); optimizer.step()
    """

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._plan = None

    def prepare(self, data_all):
        from .graph import build_full_graph
        from .surfaceNetUpdatedEdgeFilters import AttrView
        net = self.model
        dev = torch.device(net.clf.temp.device)
        n = data_all.x.shape[0]
        pos = getattr(data_all, "pos", None)
        full = build_full_graph(data_all.edge_index.to(torch.long), data_all.edge_attr, n, dev, pos=pos, order="auto",
                                need_backward=False)
        maps = build_halo_maps(full.nbr, partition_bounds(n, self.world), self.rank)
        lo, hi, n_own, n_src = maps.lo, maps.hi, maps.n_own, maps.n_own + maps.n_halo
        # local edge list: edge 4t+k runs from local_nbr[t,k] (in [owned | halo]) into owned target t
        tgt = torch.arange(n_own, device=dev).repeat_interleave(4)
        src = maps.local_nbr.reshape(-1).long()
        ei = torch.stack([src, tgt])
        e_loc = 4 * n_own
        fe_raw = data_all.edge_attr.shape[1]
        ea_loc = full.ea_in[lo:hi].reshape(e_loc, -1)[:, :fe_raw].contiguous()
        rows = torch.cat([torch.arange(lo, hi, device=dev), maps.halo_gid.to(dev)])
        x_dev = data_all.x.to(dev, dtype=torch.float32)
        x_loc = (x_dev[full.perm.long()] if full.perm is not None else x_dev).index_select(0, rows).contiguous()
        adj = (ei, torch.arange(e_loc, device=dev), (n_src, n_own))
        local = AttrView(x=x_loc, n_id=torch.arange(n_src, device=dev), adjs=[adj] * net.num_layers, edge_attr=ea_loc)
        ids = full.perm[lo:hi].long() if full.perm is not None else torch.arange(lo, hi, device=dev)
        self._plan = (local, maps, ids, HaloComm(maps, n, self.group))
        return self._plan

    def prepare_scene(self, scene):
        """Sharded build from this rank's ``dgnn_b200.scene.LocalScene`` (halo maps negotiated with the peers; the edge
        state of an edge lives with its target cell, so only ``ea_in`` is needed)."""
        from .surfaceNetUpdatedEdgeFilters import AttrView
        net = self.model
        dev = torch.device(net.clf.temp.device)
        maps = build_halo_maps_sharded(scene.nbr_gid.to(dev), scene.lo, scene.hi, partition_bounds(scene.n_global, self.world),
                                       self.rank, self.group)
        n_own, n_src = maps.n_own, maps.n_own + maps.n_halo
        comm = HaloComm(maps, scene.n_global, self.group)
        x_loc = torch.zeros((n_src, scene.x.shape[1]), dtype=torch.float32, device=dev)
        x_loc[:n_own] = scene.x
        if scene.x.shape[1] % 4 == 0:
            comm.exchange(x_loc)
        else:                                              # the row exchange moves multiples of 4 floats
            xp = torch.zeros((n_src, (scene.x.shape[1] + 3) // 4 * 4), dtype=torch.float32, device=dev)
            xp[:n_own, :scene.x.shape[1]] = scene.x
            comm.exchange(xp)
            x_loc = xp[:, :scene.x.shape[1]].contiguous()
        tgt = torch.arange(n_own, device=dev).repeat_interleave(4)
        ei = torch.stack([maps.local_nbr.reshape(-1).long(), tgt])
        e_loc = 4 * n_own
        ea_loc = scene.ea_in.reshape(e_loc, -1).contiguous()
        adj = (ei, torch.arange(e_loc, device=dev), (n_src, n_own))
        local = AttrView(x=x_loc, n_id=torch.arange(n_src, device=dev), adjs=[adj] * net.num_layers, edge_attr=ea_loc)
        self._plan = (local, maps, scene.caller_ids, comm)
        return self._plan

    def forward(self, data_all=None):
        if self._plan is None:
            self.prepare(data_all)
        local, maps, ids, comm = self._plan
        return ids, self.model(local, comm=comm)

    def allreduce_gradients(self):
        PartitionedTraining.allreduce_gradients(self)


class PartitionedInference:
    """Whole-scene inference of a ``SurfaceNet`` sharded over the ranks of the default process group.

    Every rank calls ``run(data_all)`` with the same ``data_all`` (x, edge_index, edge_attr, pos) and gets
    the logits of its owned cells plus their global (caller-order) indices."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._plan = None

    def prepare(self, data_all):
        g, maps, x0, ids, n = _local_plan(self.model, data_all, self.world, self.rank, need_backward=False)
        self._plan = (g, maps, x0, ids, HaloComm(maps, n, self.group))   # n: cells of ALL ranks (graph LayerNorm statistics)
        return self._plan

    def prepare_scene(self, scene, overlap=True):
        """Sharded build: ``scene`` is this rank's ``dgnn_b200.scene.LocalScene``; no rank holds the whole graph."""
        self._plan = _plan_from_scene(self.model, scene, self.rank, self.world, self.group, False, overlap)
        return self._plan

    @torch.no_grad()
    def run(self, data_all=None):
        from . import engine
        if self._plan is None:
            self.prepare(data_all)
        g, maps, x0, ids, comm = self._plan
        net = self.model
        with torch.cuda.device(x0.device):
            out, _ = engine.forward(net._spec(), [g] * net.num_layers, x0, training=False, save=False, comm=comm)
        return ids, out
