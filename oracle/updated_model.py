"""Plain-PyTorch CPU restatement of ``learning/surfaceNetUpdatedEdgeFilters.py``
(TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``).

Only ``SAGEConv`` (``:23-185``) and ``SurfaceNet.forward`` (``:216-251``) are meaningful
in the reference (its inference methods call the conv without ``edge_attr`` and would
raise); those are what is restated.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Linear

from .static_model import scatter_mean


class SAGEConv(nn.Module):
    """``surfaceNetUpdatedEdgeFilters.py:45-63,147-176``: returns ``(out, edge_attr')``."""

    def __init__(self, in_channels, out_channels, edge_in_channels, normalize=False, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.edge_in_channels = edge_in_channels
        self.out_channels = out_channels
        self.normalize = normalize
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        self.lin_r = Linear(in_channels, out_channels, bias=False)
        self.lin_e = Linear(edge_in_channels, in_channels, bias=bias)

    def forward(self, x, edge_attr, edge_index, size=None):
        if isinstance(x, torch.Tensor):
            x = (x, x)
        edge_attr = self.lin_e(edge_attr)  # :157
        x_j = x[0].index_select(0, edge_index[0])
        out = scatter_mean(x_j * edge_attr, edge_index[1], x[1].size(0))  # :159,:176
        out = self.lin_l(out)  # :160
        out = out + self.lin_r(x[1])  # :163-165
        if self.normalize:
            out = F.normalize(out, p=2., dim=-1)
        return out, edge_attr  # :170


class SurfaceNet(nn.Module):
    """``surfaceNetUpdatedEdgeFilters.py:189-251``."""

    def __init__(self, n_node_features, clf):
        super().__init__()
        self.clf = clf
        self.n_classes = 2
        self.n_node_feat = n_node_features
        p = clf.training.model_params
        self.convs = nn.ModuleList()
        self.convs.append(SAGEConv(self.n_node_feat, p[0], 2))  # :202
        self.convs.append(SAGEConv(p[0], p[1], self.n_node_feat))  # :203
        for i in range(len(p) - 2):  # :204-205
            self.convs.append(SAGEConv(p[i + 1], p[i + 2], p[i]))
        self.num_layers = len(self.convs)
        if clf.training.model_name[-1] == "+":  # :209-210
            self.out_net = nn.Sequential(nn.ReLU(True), nn.Linear(p[-1], 128), nn.ReLU(True),
                                         nn.Linear(128, 2))

    def forward(self, data_all):  # :216-251
        f = self.clf.features
        if f.normalization_feature and not f.keep_normalization_feature:
            x = data_all.x[data_all.n_id, 1:]
        else:
            x = data_all.x[data_all.n_id, :]
        edge_attr = data_all.edge_attr
        for i in range(self.num_layers):
            edge_index, e_id, size = data_all.adjs[i]
            new_edge_attr = torch.zeros([data_all.edge_attr.shape[0], self.convs[i].in_channels])
            x, new_edge_attr[e_id] = self.convs[i]((x, x[:size[1]]),
                                                   edge_attr[e_id, :self.convs[i].edge_in_channels],
                                                   edge_index)
            edge_attr = new_edge_attr
            if i != self.num_layers - 1:
                x = F.relu(x)
                edge_attr = F.relu(edge_attr)
        if self.clf.training.model_name[-1] == "+":
            x = F.relu(x)
            x = self.out_net(x)
        return x
