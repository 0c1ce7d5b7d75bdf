"""Full-neighbourhood closure sampler on the device.

The reference trains on seed batches: ``NeighborSampler(edge_index, node_idx, sizes=[-1] * hops, batch_size, shuffle,
return_e_id)`` (``run.py:59-74``) yields ``(batch_size, n_id, adjs)`` — for every hop all in-edges of the current node
list, the node list extended by the newly reached sources in order of first appearance, ``adjs`` outermost hop first,
each ``(edge_index in local ids, e_id, (n_src, n_tgt))`` — which ``SurfaceNet.forward`` consumes as
``data.batch_n_id / data.batch_adjs``.  PyG builds this on the host; here the k frontier expansions run on the GPU over
the in-edge ELL-4 table (``dgnn_sampler_*``), bit-exact against the oracle's restatement of the PyG sampler
(``tests/test_gpu_sampler.py``).
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch

from ._lib import DgnnError, call, check_device, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Adj(NamedTuple):
    """One hop of a sampled closure, PyG's ``Adj`` (``edge_index`` in local ids, ``e_id`` into the full edge list or
    ``None`` without ``return_e_id``, ``size = (n_src, n_tgt)``).  The reference reads ``adj.size[1]``
    (``runModel.py:274``) and ``adj.edge_index`` (``runModel.py:119,239``) as well as unpacking it as a tuple
    (``Static:215``)."""
    edge_index: torch.Tensor
    e_id: Optional[torch.Tensor]
    size: Tuple[int, int]

    def to(self, *args, **kwargs):
        return Adj(self.edge_index.to(*args, **kwargs),
                   self.e_id.to(*args, **kwargs) if self.e_id is not None else None, self.size)


class NeighborSampler:
    """Drop-in for the way the reference uses PyG's ``NeighborSampler`` (``run.py:72-74,222-223``): all neighbours
    (``sizes = [-1] * hops``), keyword arguments ``node_idx`` (index tensor or BOOL mask, as ``reduceDataset`` returns),
    ``batch_size``, ``shuffle``, ``drop_last``, ``return_e_id``, ``sampler=None``.

    ``edge_index`` int64[2, E] (``[0]`` = source, ``[1]`` = target; at most 4 in-edges per node, as in a cell graph).
    The in-edges of a node are visited in ascending (source id, edge id) order - the row order of PyG's
    ``SparseTensor(row, col, value).t()`` that ``sample_adj`` walks.  ``shuffle`` draws a new seed permutation every epoch
    with ``torch.randperm`` on the device.  Iterating yields ``(batch_size, n_id, adjs)`` with all tensors on ``device``;
    ``adjs`` is a list of ``Adj`` (outermost hop first), or a single ``Adj`` for one hop."""

    def __init__(self, edge_index: torch.Tensor, sizes, node_idx: Optional[torch.Tensor] = None,
                 num_nodes: Optional[int] = None, return_e_id=True, transform=None, batch_size: int = 1,
                 shuffle: bool = False, drop_last: bool = False, sampler=None, device="cuda:0", **loader_kwargs):
        if isinstance(sizes, int):
            sizes = [sizes]
        if any(s != -1 for s in sizes):
            raise NotImplementedError("only full neighbourhoods (sizes = [-1] * hops), as in every reference config")
        if sampler is not None or transform is not None:
            raise NotImplementedError("custom sampler / transform objects are not supported (run.py passes sampler=None)")
        unknown = set(loader_kwargs) - {"num_workers", "pin_memory", "persistent_workers", "prefetch_factor"}
        if unknown:
            raise TypeError("unexpected NeighborSampler arguments: %s" % sorted(unknown))
        self.device = torch.device(device)
        check_device(self.device.index or 0)
        self.sizes = list(sizes)
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.drop_last = bool(drop_last)
        self.return_e_id = bool(return_e_id)
        ei = edge_index.to(self.device, dtype=torch.int64).contiguous()
        if num_nodes is None and node_idx is not None and node_idx.dtype == torch.bool:
            num_nodes = node_idx.numel()                       # PyG: a mask fixes the node count
        n = (int(ei.max().item()) + 1 if ei.numel() else 0) if num_nodes is None else int(num_nodes)
        self.num_nodes = n
        # add_self_loops (run.py:70-71, edge_convs == 0): one edge i -> i per node behind the E facet edges.  The table keeps
        # the facet edges; every hop's Adj gets the self edges of its targets back (ids E + node, as add_self_loops numbers them)
        loops = ei[0] == ei[1]
        self.self_loops = bool(loops.any().item()) if ei.numel() else False
        if self.self_loops:
            if int(loops.sum().item()) != n or torch.unique(ei[0][loops]).numel() != n:
                raise DgnnError("self loops must cover every node exactly once (add_self_loops)")
            ei = ei[:, ~loops].contiguous()
        self._n_facet_edges = ei.shape[1]
        e = ei.shape[1]
        with torch.cuda.device(self.device):
            self.in_src = torch.empty((n, 4), dtype=torch.int32, device=self.device)
            self.in_eid = torch.empty((n, 4), dtype=torch.int32, device=self.device)
            cnt = torch.zeros(n, dtype=torch.int32, device=self.device)
            err = torch.zeros(1, dtype=torch.int32, device=self.device)
            call("dgnn_ell_build", ptr(ei[0]), ptr(ei[1]), e, n, n, 1, ptr(self.in_src), ptr(self.in_eid), ptr(cnt),
                 ptr(err), _stream())
            code = int(err.item())
        if code == 3:
            raise DgnnError("a node has more than 4 in-edges: not a Delaunay cell graph")
        if code:
            raise DgnnError("dgnn_ell_build: edge endpoint out of range (code %d)" % code)
        if node_idx is None:
            idx = torch.arange(n, device=self.device)
        elif node_idx.dtype == torch.bool:                     # run.py:51,72: reduceDataset's train_mask
            if node_idx.numel() != n:
                raise ValueError("node_idx mask has %d entries for %d nodes" % (node_idx.numel(), n))
            idx = node_idx.to(self.device).nonzero().view(-1)
        else:
            idx = node_idx.to(self.device, dtype=torch.int64).contiguous().view(-1)
        if idx.numel():
            if int(idx.min().item()) < 0 or int(idx.max().item()) >= n:
                raise ValueError("node_idx out of range [0, %d)" % n)
            if torch.unique(idx).numel() != idx.numel():
                raise ValueError("node_idx holds duplicate seeds")
        self.node_idx = idx
        # scratch shared by all batches: local id of a node in the current n_id (-1 = absent), first-appearance position
        self._loc = torch.full((n,), -1, dtype=torch.int32, device=self.device)
        self._first = torch.full((n,), 0x7fffffff, dtype=torch.int32, device=self.device)

    def __len__(self):
        if self.drop_last:
            return self.node_idx.numel() // self.batch_size
        return (self.node_idx.numel() + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        idx = self.node_idx
        if self.shuffle:
            idx = idx[torch.randperm(idx.numel(), device=self.device)]
        for s in range(0, idx.numel(), self.batch_size):
            batch = idx[s:s + self.batch_size]
            if self.drop_last and batch.numel() < self.batch_size:
                break
            yield self.sample(batch)

    def sample(self, batch: torch.Tensor):
        dev = self.device
        with torch.cuda.device(dev):
            st = _stream()
            n_id = batch.to(dev, dtype=torch.int64).contiguous()
            call("dgnn_sampler_set_loc", ptr(n_id), n_id.numel(), 0, ptr(self._loc), st)
            adjs = []
            for _ in self.sizes:
                n_tgt = n_id.numel()
                deg = torch.empty(n_tgt, dtype=torch.int64, device=dev)
                call("dgnn_sampler_degree", ptr(n_id), n_tgt, ptr(self.in_src), ptr(deg), st)
                incl = torch.cumsum(deg, 0)
                offs = (incl - deg).contiguous()
                n_e = int(incl[-1].item()) if n_tgt else 0
                e_id = torch.empty(n_e, dtype=torch.int64, device=dev)
                src_g = torch.empty(n_e, dtype=torch.int64, device=dev)
                edge_local = torch.empty((2, n_e), dtype=torch.int64, device=dev)
                call("dgnn_sampler_expand", ptr(n_id), n_tgt, ptr(self.in_src), ptr(self.in_eid), ptr(offs), ptr(e_id),
                     ptr(src_g), ptr(edge_local[1]), st)
                flag = torch.empty(n_e, dtype=torch.int64, device=dev)
                call("dgnn_sampler_mark", ptr(src_g), n_e, ptr(self._loc), ptr(self._first), ptr(flag), st)
                rank = torch.cumsum(flag, 0)
                n_new = int(rank[-1].item()) if n_e else 0
                new_ids = torch.empty(n_new, dtype=torch.int64, device=dev)
                call("dgnn_sampler_assign", ptr(src_g), n_e, ptr(flag), ptr(rank), n_tgt, ptr(self._loc), ptr(self._first),
                     ptr(new_ids), ptr(edge_local[0]), st)
                if self.self_loops:
                    own = torch.arange(n_tgt, dtype=torch.int64, device=dev)
                    edge_local = torch.cat([edge_local, torch.stack([own, own])], dim=1)
                    e_id = torch.cat([e_id, self._n_facet_edges + n_id[:n_tgt]])
                n_id = torch.cat([n_id, new_ids]) if n_new else n_id
                adjs.append(Adj(edge_local, e_id if self.return_e_id else None, (n_id.numel(), n_tgt)))
            call("dgnn_sampler_set_loc", ptr(n_id), n_id.numel(), -1, ptr(self._loc), st)   # scratch back to "absent"
        adjs = adjs[0] if len(adjs) == 1 else adjs[::-1]
        return batch.numel(), n_id, adjs
