"""Plain-PyTorch CPU restatement of ``learning/surfaceNetStaticEdgeFilters.py``
(TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``).

Edge-list / scatter formulation, exactly the reference's order of operations:
``lin_e`` on every edge row, ``index_select`` of the source rows, Hadamard product,
scatter-mean at the target rows, ``lin_j`` then ``+= lin_i`` (reference lines cited
inline).  Third-party semantics (PyG 2.0.2 ``MessagePassing.propagate``, torch_scatter
2.0.9 ``scatter(reduce='mean')``, PyG ``BatchNorm`` / graph-mode ``LayerNorm``) are
restated from SURVEY.md Appendix A; they are not vendored in /root/reference.

``SurfaceNet(clf)`` keeps the reference's module tree, so
``load_state_dict(kf96, strict=True)`` works.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn import Linear


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter 2.0.9 ``scatter(src, index, dim=0, dim_size, reduce='mean')``:
    scatter_add, ones-count, clamp(min=1), true divide.  Rows with no in-edge are 0."""
    out = torch.zeros((dim_size, src.size(1)), dtype=src.dtype)
    out.index_add_(0, index, src)
    cnt = torch.zeros(dim_size, dtype=src.dtype)
    cnt.index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
    cnt.clamp_(min=1)
    return out / cnt[:, None]


class BatchNorm(nn.Module):
    """PyG 2.0.2 ``nn.norm.BatchNorm``: wraps ``BatchNorm1d`` as ``.module``
    (hence the ``...norm.module.*`` state_dict keys)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, True, True)

    def forward(self, x):
        return self.module(x)


class LayerNorm(nn.Module):
    """PyG 2.0.2 ``nn.norm.LayerNorm`` with ``batch=None`` (graph mode): scalar mean /
    population std over all nodes and channels, then per-channel affine."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))

    def forward(self, x):
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + self.eps)
        return out * self.weight + self.bias


class SAGEConv(nn.Module):
    """``surfaceNetStaticEdgeFilters.py:20-109``."""

    def __init__(self, lin_i, lin_j, lin_e):
        super().__init__()
        self.lin_i = lin_i
        self.lin_j = lin_j
        self.lin_e = lin_e

    def forward(self, x, edge_attr, edge_index, size=None):
        if isinstance(x, torch.Tensor):  # :68-69
            x = (x, x)
        if self.lin_e is not None:  # :75-78
            edge_attr = self.lin_e(edge_attr)
        else:
            edge_attr = None
        # propagate (:80): x_j = x_src[edge_index[0]]; message (:89-96); mean at edge_index[1]
        x_j = x[0].index_select(0, edge_index[0])
        msg = x_j * edge_attr if edge_attr is not None else x_j
        out = scatter_mean(msg, edge_index[1], x[1].size(0))
        out = self.lin_j(out)  # :81
        out = out + self.lin_i(x[1])  # :84-86
        return out


class SurfaceNet(nn.Module):
    """``surfaceNetStaticEdgeFilters.py:114-355``."""

    def normLayer(self, size):  # :116-123
        if self.norm_type == 'b':
            return BatchNorm(size)
        elif self.norm_type == 'l':
            return LayerNorm(size)
        return None

    def sageLayer(self, inp, out):  # :125-140
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

    def __init__(self, clf):  # :146-187
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

    # -- training forward, :196-227
    def forward(self, data):
        if self.clf.regularization.cell_type:
            x = data.all.x[data.batch_n_id, 1:]
        else:
            x = data.all.x[data.batch_n_id, :]
        for i in range(self.num_layers):
            edge_index, e_id, size = data.batch_adjs[i]
            x = self.convs[i][0]((x, x[:size[1]]), data.all.edge_attr[e_id], edge_index)
            x = self.convs[i][1](x)
            x = self.convs[i][2](x)
        if self.clf.model.decoder:
            x = self.decoder(x)
        return x

    # -- whole-graph inference, :323-355
    def inference_layer(self, data_all):
        x = data_all.x[:, 1:] if self.clf.regularization.cell_type else data_all.x[:, :]
        xe = data_all.edge_attr[:, 1:] if self.clf.regularization.edge_type else data_all.edge_attr
        edge_index = data_all.edge_index.to(torch.long)
        for i in range(self.num_layers):
            x = self.convs[i][0]((x, x), xe, edge_index)
            x = self.convs[i][1](x)
            x = self.convs[i][2](x)
        if self.clf.model.decoder:
            x = self.decoder(x)
        return x

    # -- per seed batch, L-hop closure, :232-275
    def inference_batch_layer(self, data_all, batch_loader):
        x_out = torch.zeros([data_all.x.size(0), self.output_dim], dtype=torch.float32)
        x_all = data_all.x[:, 1:] if self.clf.regularization.cell_type else data_all.x
        xe = data_all.edge_attr[:, 1:] if self.clf.regularization.edge_type else data_all.edge_attr
        for batch_size, n_id, adjs in batch_loader:
            x = x_all[n_id, :]
            for i in range(self.num_layers):
                edge_index, e_id, size = adjs[i]
                x = self.convs[i][0]((x, x[:size[1]]), xe[e_id], edge_index)
                x = self.convs[i][1](x)
                x = self.convs[i][2](x)
            if self.clf.model.decoder:
                x = self.decoder(x)
            x_out[n_id[:batch_size]] = x
        return x_out

    # -- layer by layer over 1-hop batches, :279-320
    def inference_layer_batch(self, data_all, batch_loader):
        x_all = data_all.x[:, 1:] if self.clf.regularization.cell_type else data_all.x
        xe = data_all.edge_attr[:, 1:] if self.clf.regularization.edge_type else data_all.edge_attr
        for i in range(self.num_layers):
            xs = []
            for batch_size, n_id, adj in batch_loader:
                edge_index, e_id, size = adj
                x = x_all[n_id]
                x = self.convs[i][0]((x, x[:size[1]]), xe[e_id], edge_index)
                x = self.convs[i][1](x)
                x = self.convs[i][2](x)
                xs.append(x)
            x_all = torch.cat(xs, dim=0)
        return self.decoder(x_all) if self.clf.model.decoder else x_all


# --------------------------------------------------------------------------- helpers


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


class NeighborSampler:
    """Restatement of PyG 2.0.2 ``NeighborSampler(edge_index, node_idx, sizes=[-1]*h,
    batch_size, shuffle=False, return_e_id=True)`` for full neighbourhoods (SURVEY.md
    Appendix A; ``run.py:72-74,222-223``).  Yields ``(batch_size, n_id, adjs)`` with
    ``adjs`` outermost hop first; a single hop yields one tuple, not a list."""

    def __init__(self, edge_index, sizes, batch_size, node_idx=None, num_nodes=None, drop_last=False):
        self.edge_index = edge_index
        self.sizes = list(sizes)
        self.batch_size = batch_size
        self.drop_last = drop_last
        if node_idx is not None and node_idx.dtype == torch.bool:    # PyG: a mask selects its nonzero positions
            num_nodes = node_idx.numel() if num_nodes is None else num_nodes
            node_idx = node_idx.nonzero(as_tuple=False).view(-1)
        N = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
        self.N = N
        self.node_idx = torch.arange(N) if node_idx is None else node_idx
        # CSR by target: adj_t = SparseTensor(row=src, col=tgt, value=arange(E)).t() has rows = targets whose entries are
        # sorted by source id (torch_sparse sorts by row * n + col, then transposes with a stable sort by col)
        tgt = edge_index[1]
        by_src = torch.argsort(edge_index[0], stable=True)
        order = by_src[torch.argsort(tgt[by_src], stable=True)]
        self.order = order
        cnt = torch.bincount(tgt, minlength=N)
        self.rowptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(cnt, 0)])

    def __iter__(self):
        for s in range(0, self.node_idx.numel(), self.batch_size):
            b = self.node_idx[s:s + self.batch_size]
            if self.drop_last and b.numel() < self.batch_size:
                break
            yield self.sample(b)

    def __len__(self):
        if self.drop_last:
            return self.node_idx.numel() // self.batch_size
        return (self.node_idx.numel() + self.batch_size - 1) // self.batch_size

    def sample(self, batch):
        n_id = batch
        adjs = []
        for _ in self.sizes:
            n_tgt = n_id.numel()
            # all in-edges of the current targets
            starts = self.rowptr[n_id]
            ends = self.rowptr[n_id + 1]
            deg = ends - starts
            rep_t = torch.repeat_interleave(torch.arange(n_tgt), deg)
            offs = torch.arange(int(deg.sum())) - torch.repeat_interleave(torch.cumsum(deg, 0) - deg, deg)
            e_id = self.order[starts[rep_t] + offs]
            src_g = self.edge_index[0][e_id]
            # new n_id: targets first, then newly reached sources in order of first appearance
            loc = torch.full((self.N,), -1, dtype=torch.long)
            loc[n_id] = torch.arange(n_tgt)
            new_mask = loc[src_g] < 0
            new_src = src_g[new_mask]
            if new_src.numel():
                # unique preserving first appearance
                uniq, inv_idx = torch.unique(new_src, return_inverse=True)
                first = torch.full((uniq.numel(),), new_src.numel(), dtype=torch.long)
                first.scatter_reduce_(0, inv_idx, torch.arange(new_src.numel()), reduce="amin")
                uniq = uniq[torch.argsort(first)]
                loc[uniq] = n_tgt + torch.arange(uniq.numel())
                n_id = torch.cat([n_id, uniq])
            edge_local = torch.stack([loc[src_g], rep_t])
            adjs.append((edge_local, e_id, (n_id.numel(), n_tgt)))
        adjs = adjs[0] if len(adjs) == 1 else adjs[::-1]
        return batch.numel(), n_id, adjs
