"""Minimal stand-ins for the third-party modules the reference model files import at
module top, so that the reference's OWN source (``/root/reference/learning/*.py``) can be
executed unmodified in the build container to generate golden vectors.

Used ONLY by ``tests/golden/make_golden.py`` (never by the product, never on the GPU box).
What is restated here is third-party code that is NOT in /root/reference:
torch_geometric 2.0.2 ``MessagePassing`` (``propagate`` / ``__collect__`` / ``aggregate``),
``nn.norm.BatchNorm`` / ``LayerNorm`` and torch_scatter 2.0.9 ``scatter(reduce='mean')``
(SURVEY.md Appendix A).  Everything else the goldens exercise — ``SAGEConv.forward`` /
``message``, ``SurfaceNet.__init__`` / ``forward`` / ``inference_*``, ``Trainer.calcLossAndOA``
/ ``calcRegularization`` / ``train``, ``dataLoader`` — is the reference's own code.
"""
from __future__ import annotations

import inspect
import sys
import types
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn


def _scatter_mean(src, index, dim_size):
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
    cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    cnt.clamp_(min=1)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


class _Inspector:
    def __init__(self, owner):
        self.owner = owner

    def distribute(self, func_name, kwargs):
        fn = getattr(self.owner, func_name)
        names = [p for p in inspect.signature(fn).parameters if p not in ("self", "inputs")]
        return {k: kwargs[k] for k in names if k in kwargs}


class MessagePassing(nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.node_dim = 0
        self.fuse = False
        self.__explain__ = False
        self.__user_args__ = ["x_j", "edge_attr"]
        self.__fused_user_args__ = []
        self.inspector = _Inspector(self)

    def __check_input__(self, edge_index, size):
        assert isinstance(edge_index, torch.Tensor) and edge_index.dtype == torch.long
        assert edge_index.dim() == 2 and edge_index.size(0) == 2
        return [None, None] if size is None else list(size)

    def __collect__(self, args, edge_index, size, kwargs):
        out = {}
        x = kwargs["x"]
        if isinstance(x, (tuple, list)):
            size[0] = x[0].size(0) if size[0] is None else size[0]
            if x[1] is not None:
                size[1] = x[1].size(0) if size[1] is None else size[1]
            src = x[0]
        else:
            size[0] = size[1] = x.size(0)
            src = x
        out["x_j"] = src.index_select(0, edge_index[0])
        for k, v in kwargs.items():
            if k != "x":
                out[k] = v
        out["index"] = edge_index[1]
        out["ptr"] = None
        out["size"] = size
        out["dim_size"] = size[1] if size[1] is not None else size[0]
        return out

    def propagate(self, edge_index, size=None, **kwargs):
        size = self.__check_input__(edge_index, size)
        coll = self.__collect__(self.__user_args__, edge_index, size, kwargs)
        out = self.message(**self.inspector.distribute("message", coll))
        out = self.aggregate(out, **self.inspector.distribute("aggregate", coll))
        return self.update(out, **self.inspector.distribute("update", coll))

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        assert self.aggr == "mean"
        return _scatter_mean(inputs, index, dim_size)

    def update(self, inputs):
        return inputs


class BatchNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class LayerNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))

    def forward(self, x, batch=None):
        assert batch is None
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + self.eps)
        return out * self.weight + self.bias


def install():
    """Register the stand-in modules in ``sys.modules``."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    T = torch.Tensor
    mod("torch_geometric")
    mod("torch_geometric.typing", OptPairTensor=Tuple[T, Optional[T]], Adj=T,
        Size=Optional[Tuple[int, int]], OptTensor=Optional[T])
    mod("torch_geometric.nn")
    mod("torch_geometric.nn.conv", MessagePassing=MessagePassing, SAGEConv=object)
    mod("torch_geometric.nn.norm", BatchNorm=BatchNorm, LayerNorm=LayerNorm)
    mod("torch_sparse", SparseTensor=type("SparseTensor", (), {}), matmul=None)
    mod("learningHelper")
    mod("generate_mesh")  # runModel.py:16-17 (mesh extraction, off the path)
