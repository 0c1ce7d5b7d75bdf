"""The loader step right before the model, on the device.

``processing/data.py`` standardises every graph's node and edge features with a fresh sklearn ``StandardScaler``
(``standardizeFeatures``, ``data.py:467-506``: mean / population standard deviation per column in float64, constant
columns left unscaled, first column kept raw when it is the regularisation feature) and turns the adjacency file into
``edge_index`` (``readAdjacencies_bin``, ``data.py:434-439``).  ``standardize_`` does the scaler's fit + transform with
two reductions and one elementwise kernel per matrix; the statistics are float64 partial sums reduced in a fixed order.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, lib, ptr


def edge_index_from_adjacencies(adjacencies) -> torch.Tensor:
    """``readAdjacencies_bin`` (``data.py:434-439``): ``adjacencies int32[4N,2]`` (row ``4i+k`` = (cell i, its k-th facet
    neighbour)) -> ``edge_index int64[2,4N]`` with ``[0]`` = owning cell (message source), ``[1]`` = neighbour (target)."""
    a = torch.as_tensor(np.asarray(adjacencies))
    return a.t().contiguous().to(torch.int64)


def standardize_(x: torch.Tensor, skip_first: bool = False) -> torch.Tensor:
    """In-place ``StandardScaler().fit_transform`` of a float32 ``[n, c]`` device matrix (columns 1.. when
    ``skip_first``: column 0 is the raw regularisation feature, ``data.py:485-488,501-506``).  Returns ``x``."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or not x.is_contiguous():
        raise ValueError("standardize_ expects a contiguous float32 CUDA matrix")
    n, ld = x.shape
    col0 = 1 if skip_first else 0
    st = torch.cuda.current_stream().cuda_stream
    blocks = lib().dgnn_small_grid()
    for c0 in range(col0, ld, 64):
        c = min(64, ld - c0)
        part = torch.empty((blocks, 2, c), dtype=torch.float64, device=x.device)
        call("dgnn_column_moments", ptr(x), n, ld, c0, c, None, ptr(part), blocks, st)
        mean = part[:, 0].sum(0) / n
        call("dgnn_column_moments", ptr(x), n, ld, c0, c, ptr(mean), ptr(part), blocks, st)
        s = part.sum(0)
        var = (s[1] - s[0] * s[0] / n) / n               # corrected two-pass variance (population)
        scale = var.clamp_min(0).sqrt()
        # sklearn _handle_zeros_in_scale: (near-)constant columns keep their values minus the mean
        const = var <= 10 * torch.finfo(torch.float64).eps * n * mean * mean
        scale = torch.where(const | (scale == 0), torch.ones_like(scale), scale)
        call("dgnn_column_affine", ptr(x), n, ld, c0, c, ptr(mean.contiguous()), ptr((1.0 / scale).contiguous()), ld,
             ptr(x), st)
    return x
