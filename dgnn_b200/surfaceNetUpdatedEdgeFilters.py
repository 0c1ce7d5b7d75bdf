"""``learning/surfaceNetUpdatedEdgeFilters.py`` on the B200 kernels: the variant whose conv returns
``(out, edge_attr')`` so that the edge state is carried and updated from layer to layer.

Same constructor (``SurfaceNet(n_node_features, clf)``, ``Updated:191-210``), module tree and
``state_dict`` keys (``convs.{i}.lin_l/lin_r/lin_e``, ``out_net.{1,3}``) and the same
``forward(data_all)`` contract (``Updated:216-251``).  The reference's three inference methods call
the conv without ``edge_attr`` and raise; they are not provided.

Round-1 status: forward only (inference / evaluation of a trained model); per layer
  e'  = lin_e(relu?(e_prev[e_id, :edge_in]))      dgnn_layer_fwd, dense mode over the edge rows
  agg = mean_k relu?(x_src) (*) e'                 dgnn_gather_phi_fwd (edge state materialised, as the reference does)
  out = lin_l(agg) + lin_r(x_tgt)                  dgnn_dense_fwd_tc / dgnn_layer_fwd
Calling ``backward`` through it raises (no CUDA backward for this variant yet).
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


def _dense(x_in, relu_in, w, bias, n, f_in, f_out, agg=None):
    """z = [agg | relu?(x_in)] . w^T + bias on the device (tensor cores when the widths allow)."""
    dev = x_in.device
    out = torch.empty((n, f_out), dtype=torch.float32, device=dev)
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

    @torch.no_grad()
    def forward(self, data_all):
        dev = torch.device(self.clf.temp.device)
        if dev.type != "cuda":
            raise DgnnError("dgnn_b200 has no CPU path: clf.temp.device must be a CUDA (sm_100) device")
        check_device(dev.index or 0)
        if any(c.normalize for c in self.convs):
            raise NotImplementedError("normalize=True is never set by the reference model")
        with torch.cuda.device(dev):
            f = self.clf.features
            cols = slice(1, None) if (f.normalization_feature and not f.keep_normalization_feature) else slice(None)
            n_id = data_all.n_id
            x = data_all.x[n_id.to(data_all.x.device)][:, cols].to(dev, dtype=torch.float32)
            x = pad_cols(x, pad4(x.shape[1]))
            e_all = data_all.edge_attr.shape[0]
            e_state = pad_cols(data_all.edge_attr[:, :2].to(dev, dtype=torch.float32), 4)   # layer 0 reads 2 columns
            relu_in = False
            for i, conv in enumerate(self.convs):
                edge_index, e_id, size = data_all.adjs[i]
                g = build_from_edges(edge_index, None, None, size[0], size[1], dev, need_backward=False)
                n_tgt = size[1]
                fi, fo = pad4(conv.in_channels), conv.out_channels
                if fo % 4:
                    raise NotImplementedError("hidden widths must be multiples of 4")
                k_in = e_state.shape[1]
                # rows of the edge state for every (target, slot): global edge id = e_id[local edge id]
                eid_glob = torch.where(g_eid(g) >= 0, e_id.to(dev)[g_eid(g).clamp(min=0).long()].to(torch.int32),
                                       torch.full_like(g_eid(g), -1))
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
                x = _dense(x, relu_in, w_cat, conv.lin_l.bias.detach().contiguous(), n_tgt, fi, fo, agg=agg)
                # new edge state, indexed by global edge id; edges outside this hop stay 0 (Updated:236-238)
                e_state = torch.zeros((e_all, fi), dtype=torch.float32, device=dev)
                call("dgnn_scatter_rows", ptr(e_new), ptr(eid_glob), n_tgt * 4, fi, ptr(e_state), _stream())
                relu_in = True                                                        # x, e <- relu (applied on load)
            if self.clf.training.model_name[-1] == "+":
                n = x.shape[0]
                h = _dense(x, True, self.out_net[1].weight.detach(), self.out_net[1].bias.detach().contiguous(), n,
                           x.shape[1], 128)
                out = torch.empty((n, 2), dtype=torch.float32, device=dev)
                call("dgnn_rowdot_fwd", ptr(h), None, None, 1, ptr(self.out_net[3].weight.detach().contiguous()),
                     ptr(self.out_net[3].bias.detach().contiguous()), n, 128, 2, ptr(out), _stream())
                return out
            return x


def g_eid(g):
    return g._eid_in
