"""The label post-processing of ``processing/generate_mesh.py`` on the device: argmax labels
of the finite cells (``:75``), the graph-cut regularisation (``graph_cut``, ``:15-58``) and the interface facets
(``:94-105``).  Mesh assembly / metrics (trimesh, libmesh) stay with the reference.

``graph_cut(labels, prediction, edges, clf)`` keeps the reference's signature.  gco's alpha-expansion on this
two-label Potts energy ends in a global minimum, which is one s-t minimum cut; it is computed here with a lock-free
push-relabel over the ELL-4 facet table (``csrc/graphcut.cu``).  The minimum ENERGY is the reference's; where the
minimum cut is not unique (integer costs tie) the labelling may be another minimiser than the one gco happens to return.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import DgnnError, call, check_device, lib, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CutGraph:
    """Facet table of the finite cells for the cut: ``nbr int32[n,4]`` (-1 = no finite neighbour across that facet) and
    the reverse slot ``rslot uint8[n,4]`` (``nbr[nbr[c,k], rslot[c,k]] == c``), from ``edges int[E,2]`` (every
    finite-finite facet once, ``nfacets`` rows without a -1)."""

    def __init__(self, edges, n_cells: int, device):
        dev = torch.device(device)
        check_device(dev.index or 0)
        e = torch.as_tensor(np.asarray(edges)).to(dev, dtype=torch.int64).contiguous()
        ne = e.shape[0]
        self.n, self.n_edges, self.device = int(n_cells), int(ne), dev
        src = torch.cat([e[:, 0], e[:, 1]]).contiguous()
        tgt = torch.cat([e[:, 1], e[:, 0]]).contiguous()
        with torch.cuda.device(dev):
            self.nbr = torch.empty((self.n, 4), dtype=torch.int32, device=dev)
            eid = torch.empty((self.n, 4), dtype=torch.int32, device=dev)
            cnt = torch.zeros(self.n, dtype=torch.int32, device=dev)
            err = torch.zeros(1, dtype=torch.int32, device=dev)
            call("dgnn_ell_build", ptr(src), ptr(tgt), 2 * ne, self.n, self.n, 0, ptr(self.nbr), ptr(eid), ptr(cnt), ptr(err),
                 _stream())
            code = int(err.item())
            if code == 3:
                raise DgnnError("a cell has more than 4 facets: not a tetrahedralisation")
            if code:
                raise DgnnError("graph cut: facet endpoint out of range (code %d)" % code)
            # reverse slot: edge r = (a -> b) sits in row b, its twin r +- E = (b -> a) in row a
            flat = eid.reshape(-1).long()
            ok = flat >= 0
            pos = torch.zeros(max(2 * ne, 1), dtype=torch.int64, device=dev)
            pos[flat[ok]] = torch.arange(self.n * 4, device=dev)[ok]
            twin = torch.where(flat < ne, flat + ne, flat - ne).clamp(min=0)
            self.rslot = torch.where(ok, pos[twin] & 3, torch.zeros_like(flat)).to(torch.uint8).view(self.n, 4).contiguous()


def min_cut_labels(graph: CutGraph, prediction: torch.Tensor, unary_weight: float, binary_weight: int,
                   iters_per_launch: int = 32, launches_per_relabel: int = 8, return_stats: bool = False):
    """Labels (uint8, 0 = inside, 1 = outside) of a minimum of E(l) = sum_c D(c,l_c) + w * cut facets.
    ``prediction`` float32[n,2]: logits of the finite cells (NOT column-swapped; the swap of ``generate_mesh.py:25`` is
    part of the cost definition)."""
    dev = graph.device
    n = graph.n
    w = int(binary_weight)
    if w < 0:
        raise ValueError("binary_weight must be >= 0")
    z = prediction.to(dev, dtype=torch.float32).contiguous()
    if z.shape != (n, 2):
        raise ValueError("prediction must be [n_finite_cells, 2]")
    with torch.cuda.device(dev):
        st = _stream()
        excess = torch.empty(n, dtype=torch.int64, device=dev)
        sink = torch.empty(n, dtype=torch.int64, device=dev)
        call("dgnn_gc_terminals", ptr(z), n, float(unary_weight), ptr(excess), ptr(sink), None, None, st)
        cap = torch.where(graph.nbr >= 0, torch.full_like(graph.nbr, w), torch.zeros_like(graph.nbr)).contiguous()
        height = torch.empty(n, dtype=torch.int32, device=dev)
        hmax = n + 1
        changed = torch.zeros(1, dtype=torch.int32, device=dev)
        active = torch.zeros(1, dtype=torch.int64, device=dev)
        rounds = relabels = 0
        while True:
            # global relabelling: exact residual distance to the sink
            call("dgnn_gc_bfs_init", n, ptr(sink), ptr(height), hmax, st)
            level = 1
            while True:
                changed.zero_()
                for _ in range(8):                          # a few levels per host round trip
                    call("dgnn_gc_bfs_step", n, ptr(graph.nbr), ptr(graph.rslot), ptr(cap), ptr(height), level, hmax,
                         ptr(changed), st)
                    level += 1
                if int(changed.item()) == 0 or level > hmax:
                    break
            relabels += 1
            active.zero_()
            call("dgnn_gc_active", n, ptr(excess), ptr(height), hmax, ptr(active), st)
            if int(active.item()) == 0:
                break
            for _ in range(launches_per_relabel):
                call("dgnn_gc_push_relabel", n, ptr(graph.nbr), ptr(graph.rslot), ptr(cap), ptr(excess), ptr(sink), ptr(height),
                     hmax, iters_per_launch, st)
                rounds += 1
            if relabels > 100000:
                raise DgnnError("graph cut did not converge")
        labels = torch.empty(n, dtype=torch.uint8, device=dev)
        call("dgnn_gc_labels", n, ptr(height), hmax, ptr(labels), st)
    if return_stats:
        return labels, dict(push_launches=rounds, global_relabels=relabels)
    return labels


def cut_energy(graph: CutGraph, prediction: torch.Tensor, labels: torch.Tensor, unary_weight: float, binary_weight: int):
    """(data energy, smoothness energy) of a labelling, as gco's compute_data_energy / compute_smooth_energy."""
    dev = graph.device
    z = prediction.to(dev, dtype=torch.float32).contiguous()
    lab = labels.to(dev, dtype=torch.uint8).contiguous()
    with torch.cuda.device(dev):
        part = torch.zeros((lib().dgnn_gc_energy_grid(), 2), dtype=torch.int64, device=dev)
        call("dgnn_gc_energy", graph.n, ptr(z), float(unary_weight), ptr(graph.nbr), int(binary_weight), ptr(lab), ptr(part),
             _stream())
        s = part.sum(0)
    return int(s[0].item()), int(s[1].item()) // 2


def graph_cut(labels, prediction, edges, clf, device=None):
    """``processing/generate_mesh.py:graph_cut`` (same arguments: initial ``labels`` of the finite cells, their
    ``prediction`` logits, ``edges`` = finite-finite rows of ``nfacets``; ``clf.graph_cut.unary_weight / binary_weight``).
    Returns int64 NumPy labels like ``gc.get_labels()``.  The initial labels only seed gco's local search; the global
    minimum does not depend on them.  Unlike the reference, ``prediction`` is not modified in place."""
    dev = device or getattr(clf.temp, "device", None) or "cuda:0"
    pred = torch.as_tensor(prediction)
    n = int(pred.shape[0])
    edges = np.asarray(edges)
    if edges.size and int(edges.max()) + 1 > n:
        raise ValueError("edges name a cell beyond the prediction rows")
    g = CutGraph(edges, n, dev)
    lab = min_cut_labels(g, pred, float(clf.graph_cut.unary_weight), int(clf.graph_cut.binary_weight))
    return lab.cpu().numpy().astype(np.int64)


def cell_labels(prediction, infinite):
    """``generate_mesh.py:75``: labels of the finite cells, ``argmax(log_softmax(z))`` (ties -> 0), on the device."""
    from .runModel import labels as argmax_labels
    z = prediction[~infinite.to(prediction.device).bool()] if infinite is not None else prediction
    return argmax_labels(z.contiguous())


def interface_facet_ids(labels_finite, nfacets):
    """``generate_mesh.py:94-105``: indices of the facets whose two cells carry different labels, the infinite cell (-1)
    forced outside (label 1)."""
    from .runModel import interface_facets
    flag = interface_facets(labels_finite, torch.as_tensor(np.asarray(nfacets)))
    return flag.nonzero().view(-1)
