"""Drop-in replacement of ``learning/surfaceNetStaticEdgeFilters.py`` on the B200 kernels.

Same constructor (``SurfaceNet(clf)``, ``surfaceNetStaticEdgeFilters.py:146-187``), the same
module tree and therefore the same ``state_dict`` keys (``convs.{i}.conv.lin_{i,j,e}``,
``convs.{i}.norm.module.*``, ``decoder.{0,1,3}``), the same call signatures
(``forward(data)``, ``inference_layer``, ``inference_layer_batch``, ``inference_batch_layer``)
and output conventions (float32 logits on ``clf.temp.device`` in the caller's row order,
autograd-connected in training) — so ``run.py`` can import this module as ``efsage`` unchanged.

Underneath, nothing of PyG / torch_scatter is used: the graph is re-laid into ELL-4 tables
(``dgnn_b200.graph``) and every layer is one fused CUDA kernel (``dgnn_b200.engine``).  There is no
CPU path: calling the model without the built library or on a non-sm_100 device raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn import Linear

from . import engine
from ._lib import DgnnError, check_device
from .engine import ConvSpec, EdgeMlpSpec, NetSpec, NormSpec
from .graph import EllGraph, build_from_edges, build_full_graph, pad4, pad_cols


class BatchNorm(nn.Module):
    """Parameter holder with PyG's layout (``torch_geometric.nn.norm.BatchNorm`` wraps
    ``BatchNorm1d`` as ``.module``)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, True, True)

    def spec(self) -> NormSpec:
        m = self.module
        return NormSpec(0, m.weight, m.bias, m.running_mean, m.running_var, m.num_batches_tracked, m.eps, m.momentum)

    def __repr__(self):
        return "BatchNorm(%d)" % self.module.num_features


class LayerNorm(nn.Module):
    """Parameter holder with PyG's graph-mode ``LayerNorm`` layout (``weight``, ``bias``)."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))

    def spec(self) -> NormSpec:
        return NormSpec(1, self.weight, self.bias, eps=self.eps)

    def __repr__(self):
        return "LayerNorm(%d)" % self.in_channels


class SAGEConv(nn.Module):
    """Parameter holder for one edge-filtered SAGE layer (``Static:20-109``).  The arithmetic
    runs inside ``SurfaceNet`` (one fused kernel per layer, norm + ReLU included)."""

    def __init__(self, lin_i, lin_j, lin_e):
        super().__init__()
        self.lin_i = lin_i
        self.lin_j = lin_j
        self.lin_e = lin_e

    def forward(self, *a, **k):
        raise DgnnError("SAGEConv is fused into SurfaceNet's layer kernel; call the SurfaceNet instead")

    def __repr__(self):
        return '{}:\nW1: {}\nW2: {}\nΦ: {}'.format(self.__class__.__name__, self.lin_i, self.lin_j, self.lin_e)


class _NetFn(torch.autograd.Function):
    """The whole network as one autograd node: forward saves pre-norm activations and the
    aggregated messages; backward runs the hand-scheduled kernel sequence."""

    @staticmethod
    def forward(ctx, net, graphs, x0, names, comm, *params):
        spec = net._spec()
        out, sv = engine.forward(spec, graphs, x0, training=net.training, save=True, comm=comm)
        ctx.spec, ctx.sv, ctx.names, ctx.comm = spec, sv, names, comm
        return out

    @staticmethod
    def backward(ctx, dout):
        g = engine.backward(ctx.spec, ctx.sv, dout, comm=ctx.comm)
        ctx.sv = None
        return (None, None, None, None, None) + tuple(g.get(n) for n in ctx.names)


class SurfaceNet(nn.Module):

    def normLayer(self, size):  # Static:116-123
        if self.norm_type == 'b':
            return BatchNorm(size)
        elif self.norm_type == 'l':
            return LayerNorm(size)
        return None

    def sageLayer(self, inp, out):  # Static:125-140
        li = Linear(inp, out, bias=False)
        lj = Linear(inp, out, bias=True)
        if self.clf.model.edge_convs == 1:
            le = Linear(self.n_edge_feat, inp, bias=True)
        elif self.clf.model.edge_convs == 2:
            le = nn.Sequential()
            le.add_module("0", Linear(self.n_edge_feat, int(self.n_edge_feat * 2)))
            le.add_module("1", self.normLayer(int(self.n_edge_feat * 2)))
            le.add_module("2", nn.ReLU(True))
            le.add_module("3", Linear(int(self.n_edge_feat * 2), inp))
        else:
            le = None
        return SAGEConv(li, lj, le)

    def __init__(self, clf):  # Static:146-187
        super().__init__()
        self.clf = clf
        self.n_classes = 2
        self.n_node_feat = clf.temp.num_node_features
        self.n_edge_feat = clf.temp.num_edge_features
        self.norm_type = clf.model.normalization
        self.output_dim = 2 if clf.training.loss == "kl" else 1
        self.convs = nn.ModuleList()
        widths = [self.n_node_feat] + list(clf.model.convs)
        for i in range(len(widths) - 1):
            blk = nn.Sequential()
            blk.add_module("conv", self.sageLayer(widths[i], widths[i + 1]))
            blk.add_module("norm", self.normLayer(widths[i + 1]))
            blk.add_module("relu", nn.ReLU(True))
            self.convs.append(blk)
        self.num_layers = len(self.convs)
        self.decoder = nn.Sequential()
        last = clf.model.convs[-1]
        if clf.model.decoder == 1:
            self.decoder.add_module("0", nn.Linear(last, self.output_dim))
        elif clf.model.decoder == 2:
            self.decoder.add_module("0", nn.Linear(last, int(last / 2)))
            self.decoder.add_module("1", self.normLayer(int(last / 2)))
            self.decoder.add_module("2", nn.ReLU(True))
            self.decoder.add_module("3", nn.Linear(int(last / 2), self.output_dim))
        #: cache graph layouts on the data object between calls (set False to rebuild every call)
        self.cache_graphs = True
        self._pack_cache = {}     # packed kernel operands of the weights, valid until a weight changes (engine.pack_conv)

    # ------------------------------------------------------------------ parameter views
    def _spec(self) -> NetSpec:
        convs = []
        for blk in self.convs:
            c = blk.conv
            le = c.lin_e
            norm = blk.norm.spec() if blk.norm is not None else None
            if isinstance(le, nn.Sequential):        # edge_convs == 2 (Static:131-136)
                em = EdgeMlpSpec(le[0].weight, le[0].bias, le[1].spec(), le[3].weight, le[3].bias)
                convs.append(ConvSpec(c.lin_i.in_features, c.lin_i.out_features, c.lin_i.weight, c.lin_j.weight,
                                      c.lin_j.bias, None, None, norm, edge_mlp=em))
                continue
            convs.append(ConvSpec(c.lin_i.in_features, c.lin_i.out_features, c.lin_i.weight, c.lin_j.weight,
                                  c.lin_j.bias, le.weight if le is not None else None,
                                  le.bias if le is not None else None, norm))
        dec = self.clf.model.decoder or 0
        spec = NetSpec(convs, dec, out_dim=self.output_dim, cache=self._pack_cache)
        if dec == 1:
            spec.dec0_w, spec.dec0_b = self.decoder[0].weight, self.decoder[0].bias
        elif dec == 2:
            spec.dec0_w, spec.dec0_b = self.decoder[0].weight, self.decoder[0].bias
            spec.dec_norm = self.decoder[1].spec()
            spec.dec3_w, spec.dec3_b = self.decoder[3].weight, self.decoder[3].bias
        return spec

    def _named_for_grad(self):
        """(names understood by engine.backward, parameter tensors) in a fixed order."""
        names, params = [], []
        for l, blk in enumerate(self.convs):
            c = blk.conv
            names += ["convs.%d.w_i" % l, "convs.%d.w_j" % l, "convs.%d.b_j" % l]
            params += [c.lin_i.weight, c.lin_j.weight, c.lin_j.bias]
            if isinstance(c.lin_e, nn.Sequential):
                en = c.lin_e[1].spec()
                names += ["convs.%d.e0_w" % l, "convs.%d.e0_b" % l, "convs.%d.e_norm_w" % l, "convs.%d.e_norm_b" % l,
                          "convs.%d.e3_w" % l, "convs.%d.e3_b" % l]
                params += [c.lin_e[0].weight, c.lin_e[0].bias, en.weight, en.bias, c.lin_e[3].weight, c.lin_e[3].bias]
            elif c.lin_e is not None:
                names += ["convs.%d.w_e" % l, "convs.%d.b_e" % l]
                params += [c.lin_e.weight, c.lin_e.bias]
            if blk.norm is not None:
                n = blk.norm.spec()
                names += ["convs.%d.norm_w" % l, "convs.%d.norm_b" % l]
                params += [n.weight, n.bias]
        dec = self.clf.model.decoder or 0
        if dec >= 1:
            names += ["dec0_w", "dec0_b"]
            params += [self.decoder[0].weight, self.decoder[0].bias]
        if dec == 2:
            n = self.decoder[1].spec()
            names += ["dec_norm_w", "dec_norm_b", "dec3_w", "dec3_b"]
            params += [n.weight, n.bias, self.decoder[3].weight, self.decoder[3].bias]
        return tuple(names), params

    def _device(self):
        dev = torch.device(self.clf.temp.device)
        if dev.type != "cuda":
            raise DgnnError("dgnn_b200 has no CPU path: clf.temp.device must be a CUDA (sm_100) device, got %r" % (dev,))
        check_device(dev.index or 0)
        return dev

    def _run(self, graphs, x0, comm=None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            names, params = self._named_for_grad()
            return _NetFn.apply(self, graphs, x0, names, comm, *params)
        out, _ = engine.forward(self._spec(), graphs, x0, training=self.training, save=False, comm=comm)
        return out

    def _cached(self, holder, kind, tensors, build):
        """Graph layout cached on the data object.  The entry is valid only for the very same tensor OBJECTS, unmodified
        since the build (identity + ``_version`` counter); the entry keeps references to them, so an address can not be
        reused by another tensor while the entry lives."""
        key = (kind,) + tuple(None if t is None else (id(t), t._version, t.data_ptr(), tuple(t.shape)) for t in tensors)
        if self.cache_graphs:
            c = getattr(holder, "_dgnn_plan", None)
            if c is not None and c[0] == key:
                return c[1]
        plan = build()
        if self.cache_graphs:
            try:
                holder._dgnn_plan = (key, plan, list(tensors))
            except Exception:
                pass
        return plan

    # ------------------------------------------------------------------ training forward (Static:196-227)
    def forward(self, data):
        dev = self._device()
        with torch.cuda.device(dev):
            n_id = data.batch_n_id
            adjs = list(data.batch_adjs[:self.num_layers]) if isinstance(data.batch_adjs, (list, tuple)) \
                else [data.batch_adjs]
            if len(adjs) < self.num_layers:
                raise ValueError("need one adjacency per layer (%d), got %d" % (self.num_layers, len(adjs)))
            ea_all = data.all.edge_attr if self.clf.model.edge_convs else None
            full = all(a[2][0] == a[2][1] == n_id.numel() for a in adjs) and \
                all(a[0].data_ptr() == adjs[0][0].data_ptr() for a in adjs) and self.clf.model.edge_convs != 2
            keyed = [n_id, ea_all, getattr(data.all, "pos", None)] + [a[0] for a in adjs] + [a[1] for a in adjs]

            def build():
                if full:
                    ei, e_id, size = adjs[0]
                    ea = ea_all.to(dev, non_blocking=True) if ea_all is not None else None
                    pos = getattr(data.all, "pos", None)
                    pos = pos[n_id.to(pos.device)] if pos is not None else None
                    g = build_full_graph(ei, ea, size[0], dev, pos=pos, order="auto", e_id=e_id)
                    return [g] * self.num_layers
                return [build_from_edges(ei, e_id, ea_all, size[0], size[1], dev) for (ei, e_id, size) in adjs]

            graphs = self._cached(data, "train", keyed, build)
            xa = data.all.x
            cols = slice(1, None) if self.clf.regularization.cell_type else slice(None)
            if n_id.numel() * 2 >= xa.shape[0]:   # most rows used: upload once, select on the device
                x = xa.to(dev, non_blocking=True)[n_id.to(dev)][:, cols].to(torch.float32)
            else:
                x = xa[n_id.to(xa.device)][:, cols].to(dev, dtype=torch.float32, non_blocking=True)
            x0 = graphs[0].permute_rows(pad_cols(x, pad4(x.shape[1])))
            out = self._run(graphs, x0)
            return graphs[-1].unpermute_rows(out) if full else out

    # ------------------------------------------------------------------ inference (Static:232-355)
    def inference_layer(self, data_all):
        dev = self._device()
        with torch.cuda.device(dev):
            cols = slice(1, None) if self.clf.regularization.cell_type else slice(None)
            x = data_all.x[:, cols].to(dev, dtype=torch.float32, non_blocking=True)
            n = x.shape[0]
            keyed = [data_all.edge_index, data_all.edge_attr if self.clf.model.edge_convs else None,
                     getattr(data_all, "pos", None)]

            def build():
                ea = None
                if self.clf.model.edge_convs:
                    ea = data_all.edge_attr[:, 1:] if self.clf.regularization.edge_type else data_all.edge_attr
                pos = getattr(data_all, "pos", None)
                if self.clf.model.edge_convs == 2:   # the edge MLP is evaluated over the edge list
                    return build_from_edges(data_all.edge_index.to(torch.long), None, ea, n, n, dev, need_backward=False)
                return build_full_graph(data_all.edge_index.to(torch.long), ea, n, dev, pos=pos, order="auto",
                                        need_backward=False)

            g = self._cached(data_all, "infer", keyed, build)
            x0 = g.permute_rows(pad_cols(x, pad4(x.shape[1])))
            with torch.no_grad():
                out = self._run([g] * self.num_layers, x0)
            return g.unpermute_rows(out)

    def inference_layer_batch(self, data_all, batch_loader):
        """Reference: layer-by-layer over 1-hop batches with host staging (Static:279-320).
        The result is the whole-graph forward; it is computed as such on the device.  (Not so for ``normalization: l``:
        the reference's graph LayerNorm then normalises every batch by itself, Static:288-296 - no shipped config.)"""
        self._no_batched_layernorm()
        return self.inference_layer(data_all)

    def inference_batch_layer(self, data_all, batch_loader):
        """Reference: per seed batch, recompute the L-hop closure (Static:232-275).  Every
        seed's logits equal the whole-graph forward's; computed once on the device."""
        self._no_batched_layernorm()
        return self.inference_layer(data_all)

    def _no_batched_layernorm(self):
        if self.norm_type == 'l':
            raise NotImplementedError("batched inference schedules with normalization 'l': PyG's graph LayerNorm takes its "
                                      "statistics over each batch, which the whole-graph forward does not reproduce; use "
                                      "inference_layer (batch_size 0)")
