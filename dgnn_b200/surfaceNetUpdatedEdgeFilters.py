"""``learning/surfaceNetUpdatedEdgeFilters.py`` on the B200 kernels: the variant whose conv returns
``(out, edge_attr')`` so that the edge state is carried and updated from layer to layer.

Same constructor (``SurfaceNet(n_node_features, clf)``, ``Updated:191-210``), module tree and
``state_dict`` keys (``convs.{i}.lin_l/lin_r/lin_e``, ``out_net.{1,3}``) and the same
``forward(data_all)`` contract (``Updated:216-251``).  The reference's three inference methods call
the conv without ``edge_attr`` and raise; they are not provided.

Per layer (edge state materialised as ``[E, F_in]``, as the reference does, ``Updated:236``; when every layer gets the
same adjacency - a whole graph as the batch - the state stays in (target, slot) order from layer to layer and the
scatter to / gather from global edge-id order, 3 passes over ``[E, F]`` per layer and direction, is skipped):
  e'  = lin_e(relu?(e_prev[e_id, :edge_in]))      dgnn_dense_fwd_tc / dgnn_layer_fwd over the edge rows
  agg = mean_k relu?(x_src) (*) e'                 dgnn_gather_phi_fwd
  out = lin_l(agg) + lin_r(x_tgt)                  dgnn_dense_fwd_tc / dgnn_layer_fwd
Backward (autograd Function ``_UpdFn``): dgnn_dense_bwd[_tc] + dgnn_dw_bwd[_tc] for the three linears,
dgnn_upd_edge_bwd (d e', including the gradient that returns through the next layer's edge state) and
dgnn_gather_phi_bwd (d x through the out-edge table, atomic-free).  The reference's own "sage+" head cannot
run backward (in-place ReLU on a saved tensor, ``Updated:245-247``); here it can, pinned against the oracle.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn import Linear

from . import engine
from ._lib import DgnnError, call, check_device, lib, ptr
from .graph import build_from_edges, pad4, pad_cols


def _stream():
    return torch.cuda.current_stream().cuda_stream


class SAGEConv(nn.Module):
    """Parameter holder (``Updated:45-63``)."""

    def __init__(self, in_channels, out_channels, edge_in_channels, normalize=False, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.edge_in_channels = edge_in_channels
        self.out_channels = out_channels
        self.normalize = normalize
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        self.lin_r = Linear(in_channels, out_channels, bias=False)
        self.lin_e = Linear(edge_in_channels, in_channels, bias=bias)

    def __repr__(self):
        return '{}(in:{}, edge_in:{}, edge_out:{}, out:{})'.format(self.__class__.__name__, self.in_channels,
                                                                    self.edge_in_channels, self.in_channels, self.out_channels)


def _dense(x_in, relu_in, w, bias, n, f_in, f_out, agg=None, out_rows=None):
    """z = [agg | relu?(x_in)] . w^T + bias on the device (tensor cores when the widths allow).
    ``out_rows`` > n leaves room behind the result for the halo rows of a partitioned scene."""
    dev = x_in.device
    out = torch.empty((out_rows or n, f_out), dtype=torch.float32, device=dev)
    if engine.use_tensor_cores() and lib().dgnn_tc_supported(f_in, f_out, 0) and f_out % 4 == 0:
        bp = engine.pack_b(w, f_out, f_in, 2 if agg is not None else 1)
        call("dgnn_dense_fwd_tc", ptr(agg), ptr(x_in), None, None, int(relu_in), ptr(bp), ptr(bias), None, None, 0, n,
             f_in, f_out, ptr(out), None, _stream())
        return out
    if agg is not None:
        raise NotImplementedError("widths not supported by the tensor-core dense kernel")
    call("dgnn_layer_fwd", ptr(x_in), None, None, int(relu_in), None, None, 0, None, None, ptr(w.t().contiguous()),
         ptr(bias), None, None, 0, n, f_in, f_out, ptr(out), None, None, _stream())
    return out


class SurfaceNet(nn.Module):

    def __init__(self, n_node_features, clf):
        super().__init__()
        self.clf = clf
        self.n_classes = 2
        self.n_node_feat = n_node_features
        p = clf.training.model_params
        self.convs = nn.ModuleList()
        self.convs.append(SAGEConv(self.n_node_feat, p[0], 2))
        self.convs.append(SAGEConv(p[0], p[1], self.n_node_feat))
        for i in range(len(p) - 2):
            self.convs.append(SAGEConv(p[i + 1], p[i + 2], p[i], normalize=False))
        self.num_layers = len(self.convs)
        if clf.training.model_name[-1] == "+":
            self.out_net = nn.Sequential(nn.ReLU(True), nn.Linear(p[-1], 128), nn.ReLU(True), nn.Linear(128, 2))

    def _params(self):
        ps = []
        for c in self.convs:
            ps += [c.lin_l.weight, c.lin_l.bias, c.lin_r.weight, c.lin_e.weight, c.lin_e.bias]
        if self._has_head():
            ps += [self.out_net[1].weight, self.out_net[1].bias, self.out_net[3].weight, self.out_net[3].bias]
        return ps

    def _has_head(self):
        return self.clf.training.model_name[-1] == "+"

    def forward(self, data_all, comm=None):
        """``comm`` (``dgnn_b200.partition.HaloComm``, optional): ``data_all`` is one rank's [owned | halo] share of
        a partitioned scene; activations cross the partition boundary forward and gradients backward."""
        dev = torch.device(self.clf.temp.device)
        if dev.type != "cuda":
            raise DgnnError("dgnn_b200 has no CPU path: clf.temp.device must be a CUDA (sm_100) device")
        check_device(dev.index or 0)
        if any(c.normalize for c in self.convs):
            raise NotImplementedError("normalize=True is never set by the reference model")
        for c in self.convs:
            if c.out_channels % 4:
                raise NotImplementedError("hidden widths must be multiples of 4")
        params = self._params()
        if any(p.device != dev for p in params):
            raise DgnnError("model parameters are not on clf.temp.device; call model.to(device) first")
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        with torch.cuda.device(dev):
            if not need_grad:
                return _forward(self, data_all, dev, None, comm)
            return _UpdFn.apply(self, data_all, dev, comm, *params)


class _Saved:
    pass


class AttrView:
    """Minimal attribute container for a ``data_all`` built by this package (x, n_id, adjs, edge_attr)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _layer_plan(net, adj, dev, need_backward):
    """ELL tables of one hop plus the two edge-id maps the layer kernels need: ``eid_glob[t,k]`` = row of the global edge
    state behind the k-th in-edge of target t, ``orow[s,j]`` = row ``4t+k`` of e' behind the j-th out-edge of source s.
    Cached on the module per (edge_index, e_id) tensor identity + version (a whole-graph batch passes the same adjacency
    for every layer and every step; it is built once)."""
    edge_index, e_id, size = adj[0], adj[1], adj[2]
    key = (id(edge_index), edge_index._version, id(e_id), e_id._version, tuple(size), bool(need_backward), str(dev))
    cache = net.__dict__.setdefault("_plans", {})
    hit = cache.get(key)
    if hit is not None:
        return hit[0]
    g = build_from_edges(edge_index, None, None, size[0], size[1], dev, need_backward=need_backward)
    eid_loc = g._eid_in
    eid_glob = torch.where(eid_loc >= 0, e_id.to(dev)[eid_loc.clamp(min=0).long()].to(torch.int32),
                           torch.full_like(eid_loc, -1)).contiguous()
    orow = None
    if need_backward:
        row_of_edge = torch.full((max(edge_index.shape[1], 1),), -1, dtype=torch.int32, device=dev)
        flat = eid_loc.reshape(-1)
        ok = flat >= 0
        row_of_edge[flat[ok].long()] = torch.arange(flat.numel(), dtype=torch.int32, device=dev)[ok]
        eo = g._eid_out
        orow = torch.where(eo >= 0, row_of_edge[eo.clamp(min=0).long()], torch.full_like(eo, -1)).contiguous()
    # chain shortcut (see _forward): when consecutive layers share this plan and every (target, slot) carries an edge, the
    # edge state stays in incoming order from layer to layer; `ident` is the edge-id table of that order
    g._upd_all_valid = bool((eid_glob >= 0).all().item())
    g._upd_ident = torch.arange(eid_glob.numel(), dtype=torch.int32, device=dev).view_as(eid_glob) if g._upd_all_valid else None
    if len(cache) > 16:
        cache.clear()
    cache[key] = ((g, eid_glob, orow), (edge_index, e_id))      # the entry keeps the keyed tensors alive
    return g, eid_glob, orow


def _forward(net, data_all, dev, sv, comm=None):
    """Forward on the device; ``sv`` (a list) receives what the backward needs, one entry per layer."""
    f = net.clf.features
    cols = slice(1, None) if (f.normalization_feature and not f.keep_normalization_feature) else slice(None)
    n_id = data_all.n_id
    x = data_all.x[n_id.to(data_all.x.device)][:, cols].to(dev, dtype=torch.float32)
    x = pad_cols(x, pad4(x.shape[1]))
    e_all = data_all.edge_attr.shape[0]
    e_state = pad_cols(data_all.edge_attr[:, :2].to(dev, dtype=torch.float32), 4)   # layer 0 reads 2 columns
    relu_in = False
    plans = [_layer_plan(net, data_all.adjs[i], dev, sv is not None) for i in range(len(net.convs))]
    # Whole-graph batches pass the SAME adjacency to every layer (Updated:216-251 with one edge_index): then layer i + 1
    # reads, for every (target, slot), exactly the row layer i has just written for it, and the detour through the edge
    # state in global edge-id order (zero-fill + scatter here, gather there: 3 passes over [E, F] per layer, again in the
    # backward) is the identity.  `chained[i]`: layer i takes its edge rows straight from layer i - 1.
    chained = [i > 0 and plans[i][1] is plans[i - 1][1] and plans[i][0]._upd_all_valid for i in range(len(net.convs))]
    e_prev = None
    for i, conv in enumerate(net.convs):
        edge_index, e_id, size = data_all.adjs[i]
        g, eid_glob, orow = plans[i]
        n_tgt = size[1]
        fi, fo = pad4(conv.in_channels), conv.out_channels
        if chained[i]:
            ea = e_prev
            k_in = ea.shape[1]
        else:
            k_in = e_state.shape[1]
            # rows of the edge state for every (target, slot): global edge id = e_id[local edge id]
            ea = torch.empty((n_tgt * 4, k_in), dtype=torch.float32, device=dev)
            call("dgnn_gather_rows", ptr(e_state), ptr(eid_glob), n_tgt * 4, k_in, ptr(ea), _stream())
        w_e = engine._pad2(conv.lin_e.weight.detach(), fi, k_in).contiguous()
        b_e = engine._pad1(conv.lin_e.bias.detach(), fi).contiguous()
        e_new = _dense(ea, relu_in, w_e, b_e, n_tgt * 4, k_in, fi)            # pre-ReLU e' (Updated:157)
        agg = torch.empty((n_tgt, fi), dtype=torch.float32, device=dev)
        call("dgnn_gather_phi_fwd", ptr(x), None, None, int(relu_in), ptr(g.nbr), ptr(e_new), n_tgt, fi, ptr(agg),
             _stream())
        w_cat = torch.cat([engine._pad2(conv.lin_l.weight.detach(), fo, fi),
                           engine._pad2(conv.lin_r.weight.detach(), fo, fi)], dim=1).contiguous()
        halo = comm is not None and i + 1 < len(net.convs)
        out = _dense(x, relu_in, w_cat, conv.lin_l.bias.detach().contiguous(), n_tgt, fi, fo, agg=agg,
                     out_rows=size[0] if halo else None)
        if halo:
            comm.exchange(out)                 # rows n_tgt.. = the owners' outputs of the halo cells
        if sv is not None:
            s = _Saved()
            s.g, s.eid_glob, s.ea, s.phi, s.agg, s.x_in, s.out = g, eid_glob, ea, e_new, agg, x, out
            s.relu_in, s.w_e, s.w_cat, s.fi, s.fo, s.k_in, s.e_all = relu_in, w_e, w_cat, fi, fo, k_in, e_all
            s.orow = orow          # row 4t+k of e' for every out-edge (s,j)
            s.chained = chained[i]  # this layer's edge rows ARE the previous layer's e' (gradient flows back row by row)
            sv.append(s)
        x = out
        e_prev = e_new
        if i + 1 < len(net.convs) and not chained[i + 1]:
            # new edge state, indexed by global edge id; edges outside this hop stay 0 (Updated:236-238)
            e_state = torch.zeros((e_all, fi), dtype=torch.float32, device=dev)
            call("dgnn_scatter_rows", ptr(e_new), ptr(eid_glob), n_tgt * 4, fi, ptr(e_state), _stream())
        relu_in = True                                                        # x, e <- relu (applied on load)
    if net._has_head():
        n = data_all.adjs[len(net.convs) - 1][2][1]
        w1 = net.out_net[1].weight.detach().contiguous()
        h = _dense(x, True, w1, net.out_net[1].bias.detach().contiguous(), n, x.shape[1], 128)
        logits = torch.empty((n, 2), dtype=torch.float32, device=dev)
        call("dgnn_rowdot_fwd", ptr(h), None, None, 1, ptr(net.out_net[3].weight.detach().contiguous()),
             ptr(net.out_net[3].bias.detach().contiguous()), n, 128, 2, ptr(logits), _stream())
        if sv is not None:
            s = _Saved()
            s.x_in, s.h, s.w1 = x, h, w1
            sv.append(s)
        return logits
    return x


_NOAFF = engine.Affine(None, None, None, None)
_NOCOEF = (None, None, None)


class _UpdFn(torch.autograd.Function):
    """Autograd bridge: parameters enter as inputs so that ``loss.backward()`` fills their ``.grad``."""

    @staticmethod
    def forward(ctx, net, data_all, dev, comm, *params):
        sv = []
        out = _forward(net, data_all, dev, sv, comm)
        ctx.net, ctx.sv, ctx.dev, ctx.comm = net, sv, dev, comm
        return out

    @staticmethod
    def backward(ctx, dout):
        net, sv, dev, comm = ctx.net, ctx.sv, ctx.dev, ctx.comm
        st = _stream()
        grads = []
        with torch.cuda.device(dev):
            d_out = dout.contiguous().float()
            head_grads = []
            if net._has_head():
                s = sv[-1]
                n, f = s.x_in.shape
                w3 = net.out_net[3].weight.detach().contiguous()
                dy = torch.empty((n, 128), dtype=torch.float32, device=dev)
                part = torch.empty((lib().dgnn_small_grid(), 2 * 128 + 2 + 2 * 128), dtype=torch.float64, device=dev)
                call("dgnn_rowdot_bwd", ptr(d_out), ptr(s.h), None, None, None, None, 1, ptr(w3), n, 128, 2, ptr(dy),
                     ptr(part), st)
                r = engine._reduce(part)
                dw3, db3 = r[:256].view(2, 128), r[256:258]
                _, d_relu_x, db1, dw1 = engine._dense_and_dw(dy, s.h, _NOCOEF, _NOAFF, s.w1, None, None, s.x_in, None,
                                                             True, n, f, 128)
                d_out = torch.empty_like(d_relu_x)
                call("dgnn_relu_mask", ptr(d_relu_x), ptr(s.x_in), d_relu_x.numel(), ptr(d_out), st)
                head_grads = [dw1, db1, dw3, db3]
            de_next = None
            de_chained = False     # de_next is in the incoming order of the layer below (chain shortcut), not by global edge id
            layer_grads = []
            for i in range(len(net.convs) - 1, -1, -1):
                s, conv = sv[i], net.convs[i]
                g, n_tgt, n_src, fi, fo = s.g, s.g.n_tgt, s.g.n_src, s.fi, s.fo
                d_agg, d_self, db_l, dw_cat = engine._dense_and_dw(d_out, s.out, _NOCOEF, _NOAFF, s.w_cat, g, s.agg,
                                                                   s.x_in, None, s.relu_in, n_tgt, fi, fo)
                dphi = torch.empty((n_tgt * 4, fi), dtype=torch.float32, device=dev)
                call("dgnn_upd_edge_bwd", ptr(s.x_in), None, None, int(s.relu_in), ptr(g.nbr), ptr(d_agg), ptr(s.phi),
                     ptr(de_next), ptr(g._upd_ident if de_chained else s.eid_glob), n_tgt, fi, ptr(dphi), st)
                _, d_ea, db_e, dw_e = engine._dense_and_dw(dphi, s.phi, _NOCOEF, _NOAFF, s.w_e, None, None, s.ea, None,
                                                           s.relu_in, n_tgt * 4, s.k_in, fi)
                ic, ec = conv.in_channels, conv.edge_in_channels
                layer_grads.append([dw_cat[:, :ic].contiguous(), db_l, dw_cat[:, fi:fi + ic].contiguous(),
                                    dw_e[:ic, :ec].contiguous(), db_e[:ic].contiguous()])
                if i > 0:
                    d_x = torch.empty((n_src, fi), dtype=torch.float32, device=dev)
                    call("dgnn_gather_phi_bwd", ptr(d_agg), ptr(d_self), ptr(g.onbr), ptr(s.orow), ptr(s.phi),
                         ptr(s.x_in), int(s.relu_in), n_src, n_tgt, fi, ptr(d_x), st)
                    if comm is not None:
                        # the rows of halo sources hold this rank's share of their gradient: send them home, add
                        # what the peers computed for our boundary cells, and keep the owned rows
                        comm.reverse_add(d_x)
                        d_x = d_x[:sv[i - 1].g.n_tgt]
                    d_out = d_x
                    de_chained = s.chained
                    if s.chained:              # row r of this layer's edge input is row r of the layer below's e'
                        de_next = d_ea
                    else:
                        de_next = torch.zeros((s.e_all, s.k_in), dtype=torch.float32, device=dev)
                        call("dgnn_scatter_rows", ptr(d_ea), ptr(s.eid_glob), n_tgt * 4, s.k_in, ptr(de_next), st)
            for lg in reversed(layer_grads):
                grads += lg
            grads += head_grads
        ctx.sv = None
        return (None, None, None, None, *grads)


def g_eid(g):
    return g._eid_in
