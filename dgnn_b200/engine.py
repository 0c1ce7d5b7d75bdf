"""Forward / backward schedule of the Static edge-filter network over the C-ABI kernels.

Mirrors, layer by layer, what autograd does for ``SurfaceNet.forward`` /
``inference_layer`` in ``learning/surfaceNetStaticEdgeFilters.py:196-227,323-355``, but with
one fused kernel per layer (``dgnn_layer_fwd``), the norm + ReLU of layer ``l`` applied on load by
layer ``l+1``, and a hand-scheduled atomic-free backward.  Host code is PyTorch for memory,
streams and autograd plumbing only; all arithmetic on activations happens in ``libdgnn_b200.so``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from ._lib import DgnnError, call, lib, ptr
from .graph import EllGraph, pad4


def _stream():
    return torch.cuda.current_stream().cuda_stream


def use_tensor_cores() -> bool:
    """tcgen05 path for the dense transforms unless DGNN_FMA_ONLY=1 (generic FP32 path)."""
    return os.environ.get("DGNN_FMA_ONLY", "0") != "1"


def pack_b(w: torch.Tensor, n_rows: int, seg_len: int, n_segs: int, backward: bool = False) -> torch.Tensor:
    """``w`` float32[n_rows, n_segs*seg_len] (row-major) -> 128B-swizzled TF32 hi/lo K-atoms, rows grouped in the
    column slices the forward / backward launches cover."""
    w = w.contiguous()
    out = torch.empty(lib().dgnn_tc_packed_floats(n_rows, seg_len, n_segs), dtype=torch.float32, device=w.device)
    call("dgnn_pack_b_tf32", ptr(w), n_rows, w.shape[1], seg_len, n_segs, lib().dgnn_tc_slice(1 if backward else 0),
         ptr(out), _stream())
    return out


# --------------------------------------------------------------------------- parameter packing


@dataclass
class NormSpec:
    mode: int                    # 0 = BatchNorm1d, 1 = PyG graph LayerNorm
    weight: torch.Tensor
    bias: torch.Tensor
    running_mean: Optional[torch.Tensor] = None
    running_var: Optional[torch.Tensor] = None
    num_batches_tracked: Optional[torch.Tensor] = None
    eps: float = 1e-5
    momentum: float = 0.1


@dataclass
class EdgeMlpSpec:
    """``edge_convs == 2`` (Static:131-136): Linear(Fe -> 2Fe) -> norm over the layer's edges -> ReLU -> Linear(2Fe -> F_in)."""
    w0: torch.Tensor             # [2Fe, Fe]
    b0: torch.Tensor
    norm: NormSpec
    w3: torch.Tensor             # [f_in, 2Fe]
    b3: torch.Tensor


@dataclass
class ConvSpec:
    f_in: int                    # unpadded
    f_out: int
    w_i: torch.Tensor            # [f_out, f_in]
    w_j: torch.Tensor
    b_j: torch.Tensor
    w_e: Optional[torch.Tensor]  # [f_in, fe] or None (edge_convs == 0 or 2)
    b_e: Optional[torch.Tensor]
    norm: Optional[NormSpec]
    edge_mlp: Optional[EdgeMlpSpec] = None


@dataclass
class NetSpec:
    convs: List[ConvSpec]
    decoder: int                                  # 0, 1, 2
    dec0_w: Optional[torch.Tensor] = None
    dec0_b: Optional[torch.Tensor] = None
    dec_norm: Optional[NormSpec] = None
    dec3_w: Optional[torch.Tensor] = None
    dec3_b: Optional[torch.Tensor] = None
    out_dim: int = 2
    cache: Optional[dict] = None                  # packed-weight cache owned by the module (see pack_conv)


def _pad2(w: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    if w.shape == (rows, cols):
        return w
    out = torch.zeros((rows, cols), dtype=w.dtype, device=w.device)
    out[:w.shape[0], :w.shape[1]] = w
    return out


def _pad1(v: torch.Tensor, n: int) -> torch.Tensor:
    if v.shape[0] == n:
        return v
    out = torch.zeros(n, dtype=v.dtype, device=v.device)
    out[:v.shape[0]] = v
    return out


@dataclass
class PackedConv:
    f_in: int          # padded
    f_out: int
    fe: int            # padded (0 = none)
    wt_cat: torch.Tensor   # [2 f_in, f_out]
    w_cat: torch.Tensor    # [f_out, 2 f_in]
    bias: torch.Tensor
    w_e: Optional[torch.Tensor]
    b_e: Optional[torch.Tensor]
    b_fwd: Optional[torch.Tensor] = None   # tcgen05 operand of the forward  ([W_j | W_i], K-major atoms)
    b_bwd: Optional[torch.Tensor] = None   # tcgen05 operand of the backward ([W_j | W_i]^T)


def _wkey(*ts):
    return tuple(None if t is None else (id(t), t._version) for t in ts)


def pack_conv(c: ConvSpec, fe_p: int, cache: Optional[dict] = None, slot=None) -> PackedConv:
    """Concatenate / transpose / zero-pad one layer's weights into the kernel layouts.
    (Tiny glue on the weights; activations never pass through torch ops.)  With ``cache`` (a dict owned by the
    module) the packed operands are reused until a weight tensor is replaced or modified in place (tensor identity +
    autograd version counter; ``runModel.Adam`` bumps the counter after its kernel has written the parameters)."""
    key = None
    if cache is not None:
        key = ("conv", fe_p, use_tensor_cores()) + _wkey(c.w_i, c.w_j, c.b_j, c.w_e, c.b_e)
        hit = cache.get(slot)
        if hit is not None and hit[0] == key:
            return hit[1]
    fi = pad4(c.f_in)
    if c.f_out % 4:
        raise NotImplementedError("hidden widths must be multiples of 4 (got %d)" % c.f_out)
    wj = _pad2(c.w_j.detach(), c.f_out, fi)
    wi = _pad2(c.w_i.detach(), c.f_out, fi)
    w_cat = torch.cat([wj, wi], dim=1).contiguous()
    wt_cat = w_cat.t().contiguous()
    w_e = b_e = None
    if c.w_e is not None:
        w_e = _pad2(c.w_e.detach(), fi, fe_p).contiguous()
        b_e = _pad1(c.b_e.detach(), fi).contiguous()
    pk = PackedConv(fi, c.f_out, fe_p if c.w_e is not None else 0, wt_cat, w_cat, c.b_j.detach().contiguous(),
                    w_e, b_e)
    if use_tensor_cores() and lib().dgnn_tc_supported(fi, c.f_out, 1):
        pk.b_fwd = pack_b(w_cat, c.f_out, fi, 2)
    if cache is not None:
        cache[slot] = (key, pk)
    return pk


# --------------------------------------------------------------------------- normalisation helpers


@dataclass
class Affine:
    scale: torch.Tensor
    shift: torch.Tensor
    mean: Optional[torch.Tensor] = None
    rstd: Optional[torch.Tensor] = None


def eval_affine(n: NormSpec, c: int, dev) -> Affine:
    if n.mode != 0:
        raise NotImplementedError("graph LayerNorm has no running statistics; it is evaluated with batch statistics")
    scale = torch.empty(c, dtype=torch.float32, device=dev)
    shift = torch.empty(c, dtype=torch.float32, device=dev)
    call("dgnn_norm_eval_affine", ptr(n.weight), ptr(n.bias), ptr(n.running_mean), ptr(n.running_var),
         float(n.eps), c, ptr(scale), ptr(shift), _stream())
    return Affine(scale, shift)


def batch_affine(n: NormSpec, stats: torch.Tensor, n_rows: int, c: int, dev, update_running: bool) -> Affine:
    buf = torch.empty((4, c), dtype=torch.float32, device=dev)
    rm = n.running_mean if (update_running and n.mode == 0) else None
    rv = n.running_var if (update_running and n.mode == 0) else None
    call("dgnn_norm_finalize", ptr(stats), stats.shape[0], n_rows, c, ptr(n.weight), ptr(n.bias), float(n.eps),
         float(n.momentum), n.mode, ptr(rm), ptr(rv), ptr(buf[0]), ptr(buf[1]), ptr(buf[2]), ptr(buf[3]), _stream())
    if rm is not None and n.num_batches_tracked is not None:
        n.num_batches_tracked += 1
    return Affine(buf[0], buf[1], buf[2], buf[3])


# --------------------------------------------------------------------------- forward


@dataclass
class Saved:
    x0: torch.Tensor = None
    z: List[torch.Tensor] = field(default_factory=list)        # pre-norm outputs per conv layer
    agg: List[torch.Tensor] = field(default_factory=list)
    aff: List[Optional[Affine]] = field(default_factory=list)  # norm affine per conv layer
    packed: List[PackedConv] = field(default_factory=list)
    z_d: Optional[torch.Tensor] = None
    aff_d: Optional[Affine] = None
    graphs: List[EllGraph] = field(default_factory=list)
    edge: List[Optional[tuple]] = field(default_factory=list)   # edge_convs == 2: (e1, aff_e, phi_in, w0p, w3p) per layer


def _layer_fwd(x_in, in_aff: Optional[Affine], relu_in: bool, g: Optional[EllGraph], pk_wt, pk_bias, w_e, b_e, fe,
               out_aff: Optional[Affine], relu_out: bool, n_tgt, f_in, f_out, want_agg, want_stats, b_packed=None,
               out_rows=None):
    dev = x_in.device
    loops = g is not None and g.self_loops
    if loops:
        # add_self_loops (run.py:70-71, edge_convs == 0): the generic kernel takes the flag as bit 1 of relu_in
        if fe:
            raise DgnnError("self loops are only defined without an edge filter (model.edge_convs == 0)")
        b_packed = None
    out = torch.empty((out_rows or n_tgt, f_out), dtype=torch.float32, device=dev)
    agg = torch.empty((n_tgt, f_in), dtype=torch.float32, device=dev) if want_agg else None
    stats = None
    if want_stats:
        grid = lib().dgnn_tc_grid() if b_packed is not None else lib().dgnn_layer_grid(f_in, f_out)
        stats = torch.empty((grid, 2, f_out), dtype=torch.float64, device=dev)
    call("dgnn_layer_fwd_tc" if b_packed is not None else "dgnn_layer_fwd", ptr(x_in), ptr(in_aff.scale) if in_aff else None, ptr(in_aff.shift) if in_aff else None,
         int(relu_in) | (2 if loops else 0), ptr(g.nbr) if g is not None else None,
         ptr(g.ea_in) if (g is not None and fe) else None, fe,
         ptr(w_e), ptr(b_e), ptr(b_packed) if b_packed is not None else ptr(pk_wt), ptr(pk_bias),
         ptr(out_aff.scale) if out_aff else None, ptr(out_aff.shift) if out_aff else None, int(relu_out),
         n_tgt, f_in, f_out, ptr(out), ptr(agg), ptr(stats), _stream())
    return out, agg, stats


#: SMs left free for the pack + NCCL kernels while a halo exchange overlaps a layer's interior rows
OVERLAP_RESERVED_SMS = 8


def _gather_then_dense(x_in, in_aff, relu_in, g: EllGraph, pk: PackedConv, out_aff, relu_out, want_stats,
                       out_rows=None, comm=None):
    """Tensor-core layer forward as two kernels: aggregation with the edge filter on tcgen05
    (dgnn_gather_tc_fwd) -> agg, then z = [agg | h] . W^T (dgnn_dense_fwd_tc).

    With a boundary-first partition (``comm.n_boundary > 0``) the target rows are processed as two ranges: the
    boundary rows first, then - while their halo exchange runs on the communication stream (``comm.start``) - the
    interior rows, with a few SMs left free for the exchange.  Returns ``(out, agg, stats, exchange_started)``."""
    dev = x_in.device
    n_tgt = g.n_tgt
    f_in, f_out, fe = pk.f_in, pk.f_out, pk.fe
    agg = torch.empty((n_tgt, f_in), dtype=torch.float32, device=dev)
    out = torch.empty((out_rows or n_tgt, f_out), dtype=torch.float32, device=dev)
    sc = ptr(in_aff.scale) if in_aff else None
    sh = ptr(in_aff.shift) if in_aff else None
    nb = comm.n_boundary if (comm is not None and out_rows is not None and out_rows > n_tgt) else 0
    ranges = [(0, nb), (nb, n_tgt)] if 0 < nb < n_tgt else [(0, n_tgt)]
    grid = lib().dgnn_tc_grid()
    stats = torch.zeros((len(ranges) * grid, 2, f_out), dtype=torch.float64, device=dev) if want_stats else None
    started = False
    for i, (r0, r1) in enumerate(ranges):
        prev = lib().dgnn_reserve_sms(OVERLAP_RESERVED_SMS) if (i == 1) else None
        try:
            call("dgnn_gather_tc_fwd", ptr(x_in), sc, sh, int(relu_in), ptr(g.nbr) + r0 * 16, ptr(g.ea_in) + r0 * 16 * fe, fe,
                 ptr(pk.w_e), ptr(pk.b_e), r1 - r0, f_in, ptr(agg) + r0 * 4 * f_in, _stream())
            call("dgnn_dense_fwd_tc", ptr(agg) + r0 * 4 * f_in, ptr(x_in) + r0 * 4 * f_in, sc, sh, int(relu_in), ptr(pk.b_fwd),
                 ptr(pk.bias), ptr(out_aff.scale) if out_aff else None, ptr(out_aff.shift) if out_aff else None,
                 int(relu_out), r1 - r0, f_in, f_out, ptr(out) + r0 * 4 * f_out,
                 ptr(stats) + i * grid * 2 * f_out * 8 if want_stats else None, _stream())
        finally:
            if prev is not None:
                lib().dgnn_reserve_sms(prev)
        if i == 0 and len(ranges) == 2:
            comm.start(out)                     # boundary rows are final: pack + all-to-all on the communication stream
            started = True
    return out, agg, stats, started


def _dense_from_agg(agg, x_in, in_aff, relu_in, n_tgt, pk: PackedConv, out_aff, relu_out, want_stats, out_rows=None):
    """z = [agg | h(x_in)] . [W_j | W_i]^T (dgnn_dense_fwd_tc)."""
    dev = x_in.device
    sc = ptr(in_aff.scale) if in_aff else None
    sh = ptr(in_aff.shift) if in_aff else None
    out = torch.empty((out_rows or n_tgt, pk.f_out), dtype=torch.float32, device=dev)
    stats = torch.empty((lib().dgnn_tc_grid(), 2, pk.f_out), dtype=torch.float64, device=dev) if want_stats else None
    call("dgnn_dense_fwd_tc", ptr(agg), ptr(x_in), sc, sh, int(relu_in), ptr(pk.b_fwd), ptr(pk.bias),
         ptr(out_aff.scale) if out_aff else None, ptr(out_aff.shift) if out_aff else None, int(relu_out), n_tgt,
         pk.f_in, pk.f_out, ptr(out), ptr(stats), _stream())
    return out, agg, stats


def _edge_mlp_fwd(em: EdgeMlpSpec, g: EllGraph, f_in_p: int, batch_stats: bool, training: bool):
    """phi for every (target, slot) from the two-layer edge MLP, evaluated over the layer's edges in edge-list order so
    that the norm statistics are taken over exactly the rows the reference normalises (``edge_attr[e_id]``).
    Returns (e1 pre-norm [E, 2Fe], its norm affine, phi [n_tgt*4, f_in] in incoming order, padded weights)."""
    if g.ea_edges is None or g._eid_in is None:
        raise DgnnError("edge_convs == 2 needs the edge-list layout (graph.build_from_edges)")
    ea = g.ea_edges
    dev = ea.device
    n_e, fe_p = ea.shape
    h1 = em.w0.shape[0]
    if h1 % 4:
        raise NotImplementedError("edge MLP hidden width must be a multiple of 4")
    w0p = _pad2(em.w0.detach(), h1, fe_p).contiguous()
    w3p = _pad2(em.w3.detach(), f_in_p, h1).contiguous()
    b3p = _pad1(em.b3.detach(), f_in_p).contiguous()
    if batch_stats or em.norm.mode == 1:
        e1, _, stats = _layer_fwd(ea, None, False, None, w0p.t().contiguous(), em.b0.detach().contiguous(), None, None, 0,
                                  None, False, n_e, fe_p, h1, False, True)
        aff_e = batch_affine(em.norm, stats, n_e, h1, dev, update_running=training)
    else:
        e1, _, _ = _layer_fwd(ea, None, False, None, w0p.t().contiguous(), em.b0.detach().contiguous(), None, None, 0,
                              None, False, n_e, fe_p, h1, False, False)
        aff_e = eval_affine(em.norm, h1, dev)
    phi_e, _, _ = _layer_fwd(e1, aff_e, True, None, w3p.t().contiguous(), b3p, None, None, 0, None, False, n_e, h1,
                             f_in_p, False, False)
    phi_in = torch.empty((g.n_tgt * 4, f_in_p), dtype=torch.float32, device=dev)
    call("dgnn_gather_rows", ptr(phi_e), ptr(g._eid_in), g.n_tgt * 4, f_in_p, ptr(phi_in), _stream())
    return e1, aff_e, phi_in, w0p, w3p


def _layer_fwd_edge_mlp(h, in_aff, relu_in, g, pk, phi_in, out_aff, relu_out, want_stats, out_rows=None):
    """One conv layer whose edge filter is a materialised matrix: aggregation (dgnn_gather_phi_fwd), then the dense part."""
    if pk.b_fwd is None:
        raise NotImplementedError("edge_convs == 2 needs widths the tensor-core dense kernel supports")
    agg = torch.empty((g.n_tgt, pk.f_in), dtype=torch.float32, device=h.device)
    call("dgnn_gather_phi_fwd", ptr(h), ptr(in_aff.scale) if in_aff else None, ptr(in_aff.shift) if in_aff else None,
         int(relu_in), ptr(g.nbr), ptr(phi_in), g.n_tgt, pk.f_in, ptr(agg), _stream())
    return _dense_from_agg(agg, h, in_aff, relu_in, g.n_tgt, pk, out_aff, relu_out, want_stats, out_rows)


def _global_stats(stats, n_rows, comm):
    """Batch statistics over all ranks of a partitioned scene: sum the per-CTA partials, all-reduce the
    2 x C doubles, and count the rows of every rank."""
    if comm is None:
        return stats, n_rows
    s = stats.sum(dim=0, keepdim=True).contiguous()
    comm.allreduce(s)
    return s, comm.n_rows


def forward(spec: NetSpec, graphs: List[EllGraph], x0: torch.Tensor, training: bool, save: bool, comm=None):
    """Run all conv layers + decoder.  ``x0``: float32[n_src0, pad4(F0)] in the graphs' row order.
    ``comm`` (partitioned scenes, ``dgnn_b200.partition.HaloComm``): ``comm.exchange(h)`` fills the halo rows
    ``h[n_tgt:]`` of a layer's output in place (outputs are then allocated with ``n_src`` rows),
    ``comm.allreduce(t)`` sums a small tensor over the ranks (batch statistics), ``comm.n_rows`` = cells of
    all ranks.
    Returns ``(logits, Saved or None)``.  In training mode the norms use batch statistics
    (and update the running buffers); in eval mode the running-statistic affine + ReLU is fused
    into each layer's epilogue."""
    dev = x0.device
    sv = Saved() if save else None
    fe_p = graphs[0].fe
    h = x0
    in_aff: Optional[Affine] = None
    relu_in = False
    if save:
        sv.x0 = x0
        sv.graphs = graphs
    L = len(spec.convs)
    batch_stats = training or any(c.norm is not None and c.norm.mode == 1 for c in spec.convs)
    for l, c in enumerate(spec.convs):
        g = graphs[l]
        pk = pack_conv(c, fe_p, spec.cache, l)
        if h.shape[1] != pk.f_in:
            raise ValueError("layer %d expects %d input features, got %d" % (l, pk.f_in, h.shape[1]))
        if c.norm is None:
            raise NotImplementedError("normalization must be 'b' or 'l' (the reference crashes otherwise, Static:218)")
        split = pk.b_fwd is not None and pk.fe > 0 and lib().dgnn_gather_tc_supported(pk.f_in, pk.fe)
        out_rows = g.n_src if (comm is not None and l + 1 < L) else g.n_tgt
        edge = None
        if c.edge_mlp is not None:
            if comm is not None:
                raise NotImplementedError("edge_convs == 2 on a partitioned scene")
            edge = _edge_mlp_fwd(c.edge_mlp, g, pk.f_in, batch_stats, training)
        if save:
            sv.edge.append(edge)
        started = False
        if batch_stats:
            if edge is not None:
                z, agg, stats = _layer_fwd_edge_mlp(h, in_aff, relu_in, g, pk, edge[2], None, False, True, out_rows)
            elif split:
                z, agg, stats, started = _gather_then_dense(h, in_aff, relu_in, g, pk, None, False, True,
                                                            out_rows=out_rows, comm=comm)
            else:
                z, agg, stats = _layer_fwd(h, in_aff, relu_in, g, pk.wt_cat, pk.bias, pk.w_e, pk.b_e, pk.fe, None,
                                           False, g.n_tgt, pk.f_in, pk.f_out, save, True, b_packed=pk.b_fwd,
                                           out_rows=out_rows)
            stats, n_rows = _global_stats(stats, g.n_tgt, comm)
            aff = batch_affine(c.norm, stats, n_rows, pk.f_out, dev, update_running=training)
            if comm is not None and l + 1 < L:         # halo rows carry the owners' pre-norm z; the affine is global
                comm.finish() if started else comm.exchange(z)
            if save:
                sv.z.append(z); sv.agg.append(agg); sv.aff.append(aff); sv.packed.append(pk)
            h, in_aff, relu_in = z, aff, True
        else:
            aff = eval_affine(c.norm, pk.f_out, dev)
            if edge is not None:
                h, _, _ = _layer_fwd_edge_mlp(h, in_aff, relu_in, g, pk, edge[2], aff, True, False, out_rows)
            elif split:
                h, _, _, started = _gather_then_dense(h, in_aff, relu_in, g, pk, aff, True, False, out_rows=out_rows,
                                                      comm=comm)
            else:
                h, _, _ = _layer_fwd(h, in_aff, relu_in, g, pk.wt_cat, pk.bias, pk.w_e, pk.b_e, pk.fe, aff, True,
                                     g.n_tgt, pk.f_in, pk.f_out, False, False, b_packed=pk.b_fwd, out_rows=out_rows)
            if comm is not None and l + 1 < L:
                comm.finish() if started else comm.exchange(h)
            in_aff, relu_in = None, False
    n_out = graphs[-1].n_tgt
    f_last = spec.convs[-1].f_out
    if spec.decoder == 0:
        out = torch.empty((n_out, f_last), dtype=torch.float32, device=dev)
        call("dgnn_affine_relu", ptr(h), ptr(in_aff.scale) if in_aff else None, ptr(in_aff.shift) if in_aff else None,
             int(relu_in), n_out, f_last, ptr(out), _stream())
        return out, sv
    if spec.decoder == 1:
        w, b = spec.dec0_w.detach().contiguous(), spec.dec0_b.detach().contiguous()
        out = torch.empty((n_out, spec.out_dim), dtype=torch.float32, device=dev)
        call("dgnn_rowdot_fwd", ptr(h), ptr(in_aff.scale) if in_aff else None, ptr(in_aff.shift) if in_aff else None,
             int(relu_in), ptr(w), ptr(b), n_out, f_last, spec.out_dim, ptr(out), _stream())
        return out, sv
    # decoder == 2: Linear -> norm -> ReLU -> Linear
    f_d = spec.dec0_w.shape[0]
    if f_d % 4:
        raise NotImplementedError("decoder hidden width must be a multiple of 4")
    wt = spec.dec0_w.detach().t().contiguous()
    b0 = spec.dec0_b.detach().contiguous()
    dn = spec.dec_norm
    bd = None
    if use_tensor_cores() and lib().dgnn_tc_supported(f_last, f_d, 0):
        dkey = ("dec",) + _wkey(spec.dec0_w)
        hit = spec.cache.get("dec") if spec.cache is not None else None
        if hit is not None and hit[0] == dkey:
            bd = hit[1]
        else:
            bd = pack_b(spec.dec0_w.detach(), f_d, f_last, 1)
            if spec.cache is not None:
                spec.cache["dec"] = (dkey, bd)
    if batch_stats:
        z_d, _, stats = _layer_fwd(h, in_aff, relu_in, None, wt, b0, None, None, 0, None, False, n_out, f_last, f_d,
                                   False, True, b_packed=bd)
        stats, n_rows = _global_stats(stats, n_out, comm)
        aff_d = batch_affine(dn, stats, n_rows, f_d, dev, update_running=training)
        hd, hd_aff, hd_relu = z_d, aff_d, True
        if save:
            sv.z_d, sv.aff_d = z_d, aff_d
    else:
        aff_d = eval_affine(dn, f_d, dev)
        hd, _, _ = _layer_fwd(h, in_aff, relu_in, None, wt, b0, None, None, 0, aff_d, True, n_out, f_last, f_d,
                              False, False, b_packed=bd)
        hd_aff, hd_relu = None, False
    w3, b3 = spec.dec3_w.detach().contiguous(), spec.dec3_b.detach().contiguous()
    out = torch.empty((n_out, spec.out_dim), dtype=torch.float32, device=dev)
    call("dgnn_rowdot_fwd", ptr(hd), ptr(hd_aff.scale) if hd_aff else None, ptr(hd_aff.shift) if hd_aff else None,
         int(hd_relu), ptr(w3), ptr(b3), n_out, f_d, spec.out_dim, ptr(out), _stream())
    return out, sv


# --------------------------------------------------------------------------- backward


#: set to a dict to record backward intermediates (development / tests only)
DEBUG = None


def _dbg(name, t):
    if DEBUG is not None:
        DEBUG[name] = t.detach().clone() if t is not None else None


def _reduce(partials: torch.Tensor) -> torch.Tensor:
    """Sum double partials [P, len] -> float32[len]."""
    out = torch.empty(partials.shape[1], dtype=torch.float32, device=partials.device)
    call("dgnn_reduce_partials", ptr(partials), partials.shape[0], partials.shape[1], ptr(out), _stream())
    return out


def _norm_coeffs(n: NormSpec, aff: Affine, s1, s2, n_rows, c):
    buf = torch.empty((3, c), dtype=torch.float32, device=s1.device)
    call("dgnn_norm_bwd_coeffs", ptr(s1), ptr(s2), n_rows, c, ptr(n.weight), ptr(aff.rstd), n.mode,
         ptr(buf[0]), ptr(buf[1]), ptr(buf[2]), _stream())
    return buf[0], buf[1], buf[2]


def _dense_and_dw(dy, z, coeffs, aff, w_cat, g: Optional[EllGraph], agg, x_in, in_aff, relu_in, n_tgt, f_in, f_out,
                  agg_rows=None):
    """dz . W (-> d_agg, d_self, db) and dz^T . [agg | h] (-> dW_cat).  ``agg_rows`` > n_tgt leaves room
    behind d_agg for the halo rows of a partitioned scene."""
    dev = dy.device
    k_total = 2 * f_in if g is not None else f_in
    gq, aq, bq = coeffs
    d_self = torch.empty((n_tgt, f_in), dtype=torch.float32, device=dev)
    d_agg = torch.empty((agg_rows or n_tgt, f_in), dtype=torch.float32, device=dev) if g is not None else None
    tc_dw = use_tensor_cores() and lib().dgnn_dw_tc_supported(f_out, k_total)
    tc_dense = use_tensor_cores() and lib().dgnn_tc_supported(f_in, f_out, 1 if g is not None else 0)
    db_from_dw = tc_dw and tc_dense          # the dW kernel forms dz anyway and sums its columns (db) on the way
    if tc_dense and not db_from_dw and f_out > 256:
        tc_dense = False                     # the dense backward's shared column sums hold 256 channels
    if tc_dense:
        # operand B of the backward: [W_j | W_i]^T, i.e. rows = columns of d[agg|self], K = f_out
        b_bwd = pack_b(w_cat.t().contiguous(), k_total, f_out, 1, backward=True)
        db_p = torch.empty((lib().dgnn_tc_grid(), f_out), dtype=torch.float64, device=dev)
        call("dgnn_dense_bwd_tc", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(aff.mean), ptr(aff.rstd), ptr(b_bwd),
             ptr(g.nbr) if g is not None else None, n_tgt, f_in, f_out, ptr(d_agg), ptr(d_self),
             None if db_from_dw else ptr(db_p), _stream())
    else:
        grid = lib().dgnn_layer_grid(f_in, f_out)
        db_p = torch.empty((grid, f_out), dtype=torch.float64, device=dev)
        call("dgnn_dense_bwd", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(aff.mean), ptr(aff.rstd), ptr(w_cat),
             ptr(g.nbr) if g is not None else None, n_tgt, f_in, f_out, k_total, ptr(d_agg), ptr(d_self), ptr(db_p),
             _stream())
    splits = lib().dgnn_tc_grid() if tc_dw else lib().dgnn_dw_splits(f_out, k_total)
    dw_p = torch.empty((splits, f_out, k_total), dtype=torch.float32, device=dev)
    args = (ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(aff.mean), ptr(aff.rstd), ptr(agg), ptr(x_in),
            ptr(in_aff.scale) if in_aff else None, ptr(in_aff.shift) if in_aff else None, int(relu_in), n_tgt, f_in,
            f_out, k_total, ptr(dw_p))
    if tc_dw:
        call("dgnn_dw_bwd_tc", *args, ptr(db_p) if db_from_dw else None, _stream())
    else:
        call("dgnn_dw_bwd", *args, _stream())
    db = _reduce(db_p)
    dw = torch.empty((f_out, k_total), dtype=torch.float32, device=dev)
    call("dgnn_reduce_partials_f32", ptr(dw_p), splits, f_out * k_total, ptr(dw), _stream())
    return d_agg, d_self, db, dw


def _global_sums(s1, s2, comm):
    """(S1, S2) of a norm backward over all ranks (the local ones stay the norm's own weight/bias gradient)."""
    if comm is None:
        return s1, s2
    both = torch.cat([s1, s2])
    comm.allreduce(both)
    return both[:s1.numel()], both[s1.numel():]


def backward(spec: NetSpec, sv: Saved, dout: torch.Tensor, comm=None):
    """Gradients of every parameter given ``dout`` = dL/d(output of ``forward``).
    Returns a dict name -> grad keyed like ``NetSpec`` fields (``convs.{l}.w_i`` ...).
    With ``comm`` (partitioned scene) the returned gradients are this rank's share (sum them over ranks);
    the norm statistics are all-reduced inside and d_agg crosses the partition boundary through the same
    halo exchange as the forward activations."""
    dev = dout.device
    st = _stream()
    grads = {}
    L = len(spec.convs)
    n_out = sv.graphs[-1].n_tgt
    f_last = spec.convs[-1].f_out
    zL, affL = sv.z[-1], sv.aff[-1]
    small = lib().dgnn_small_grid()
    dout = dout.contiguous()
    # ---- decoder
    if spec.decoder == 0:
        dy = torch.empty((n_out, f_last), dtype=torch.float32, device=dev)
        part = torch.empty((small, 2 * f_last), dtype=torch.float64, device=dev)
        call("dgnn_act_bwd", ptr(dout), ptr(zL), ptr(affL.scale), ptr(affL.shift), ptr(affL.mean), ptr(affL.rstd), 1,
             n_out, f_last, ptr(dy), ptr(part), st)
        r = _reduce(part)
        s1, s2 = r[:f_last], r[f_last:]
    elif spec.decoder == 1:
        od = spec.out_dim
        dy = torch.empty((n_out, f_last), dtype=torch.float32, device=dev)
        part = torch.empty((small, od * f_last + od + 2 * f_last), dtype=torch.float64, device=dev)
        call("dgnn_rowdot_bwd", ptr(dout), ptr(zL), ptr(affL.scale), ptr(affL.shift), ptr(affL.mean), ptr(affL.rstd), 1,
             ptr(spec.dec0_w.detach().contiguous()), n_out, f_last, od, ptr(dy), ptr(part), st)
        r = _reduce(part)
        grads["dec0_w"] = r[:od * f_last].view(od, f_last)
        grads["dec0_b"] = r[od * f_last:od * f_last + od]
        s1, s2 = r[od * f_last + od:od * f_last + od + f_last], r[od * f_last + od + f_last:]
    else:
        od = spec.out_dim
        f_d = spec.dec0_w.shape[0]
        dy_d = torch.empty((n_out, f_d), dtype=torch.float32, device=dev)
        part = torch.empty((small, od * f_d + od + 2 * f_d), dtype=torch.float64, device=dev)
        affd = sv.aff_d
        call("dgnn_rowdot_bwd", ptr(dout), ptr(sv.z_d), ptr(affd.scale), ptr(affd.shift), ptr(affd.mean), ptr(affd.rstd),
             1, ptr(spec.dec3_w.detach().contiguous()), n_out, f_d, od, ptr(dy_d), ptr(part), st)
        r = _reduce(part)
        grads["dec3_w"] = r[:od * f_d].view(od, f_d)
        grads["dec3_b"] = r[od * f_d:od * f_d + od]
        s1d, s2d = r[od * f_d + od:od * f_d + od + f_d], r[od * f_d + od + f_d:]
        grads["dec_norm_w"], grads["dec_norm_b"] = s2d, s1d
        _dbg("dy_d", dy_d)
        s1g, s2g = _global_sums(s1d, s2d, comm)
        coeffs = _norm_coeffs(spec.dec_norm, affd, s1g, s2g, comm.n_rows if comm else n_out, f_d)
        _, dh, db0, dw0 = _dense_and_dw(dy_d, sv.z_d, coeffs, affd, spec.dec0_w.detach().contiguous(), None, None, zL,
                                        affL, True, n_out, f_last, f_d)
        grads["dec0_w"], grads["dec0_b"] = dw0, db0
        _dbg("dh_L", dh)
        dy = dh  # in place: dy = relu'(y_L) * dh
        part = torch.empty((small, 2 * f_last), dtype=torch.float64, device=dev)
        call("dgnn_act_bwd", ptr(dh), ptr(zL), ptr(affL.scale), ptr(affL.shift), ptr(affL.mean), ptr(affL.rstd), 1,
             n_out, f_last, ptr(dy), ptr(part), st)
        r = _reduce(part)
        s1, s2 = r[:f_last], r[f_last:]
    # ---- conv layers, last to first
    for l in range(L - 1, -1, -1):
        c, pk, g, aff = spec.convs[l], sv.packed[l], sv.graphs[l], sv.aff[l]
        grads["convs.%d.norm_w" % l], grads["convs.%d.norm_b" % l] = s2, s1
        _dbg("dy_%d" % l, dy)
        s1g, s2g = _global_sums(s1, s2, comm)
        coeffs = _norm_coeffs(c.norm, aff, s1g, s2g, comm.n_rows if comm else g.n_tgt, pk.f_out)
        x_in = sv.z[l - 1] if l > 0 else sv.x0
        in_aff = sv.aff[l - 1] if l > 0 else None
        relu_in = l > 0
        d_agg, d_self, db, dw = _dense_and_dw(dy, sv.z[l], coeffs, aff, pk.w_cat, g, sv.agg[l], x_in, in_aff, relu_in,
                                              g.n_tgt, pk.f_in, pk.f_out, agg_rows=g.n_src if comm else None)
        if g.self_loops:
            # the self edge t -> t: the mean runs over cnt + 1 rows (the kernel scaled d_agg by 1 / max(cnt, 1)), and the
            # cell's own activation receives d_agg[t] like any other source of t
            if comm is not None:
                raise DgnnError("self loops on a partitioned scene are not supported")
            cnt = (g.nbr >= 0).sum(1, dtype=torch.float32)
            d_agg[:g.n_tgt] *= (cnt.clamp(min=1.0) / (cnt + 1.0)).unsqueeze(1)
            d_self += d_agg[:g.n_tgt]
        # sources whose gradient this rank produces: all of them, or (partitioned) the owned rows only, with the
        # d_agg rows of halo targets fetched from their owners
        n_srcs = g.n_src
        if comm is not None:
            comm.exchange(d_agg)
            n_srcs = g.n_tgt
        grads["convs.%d.b_j" % l] = db
        _dbg("d_agg_%d" % l, d_agg); _dbg("d_self_%d" % l, d_self)
        grads["convs.%d.w_j" % l] = dw[:, :c.f_in]
        grads["convs.%d.w_i" % l] = dw[:, pk.f_in:pk.f_in + c.f_in]
        need_prev = l > 0
        tc_gather = use_tensor_cores() and pk.fe > 0 and lib().dgnn_gather_tc_supported(pk.f_in, pk.fe)
        if c.edge_mlp is not None:
            em = c.edge_mlp
            e1, aff_e, phi_in, w0p, w3p = sv.edge[l]
            n_e, h1 = e1.shape
            sc_in = ptr(in_aff.scale) if in_aff else None
            sh_in = ptr(in_aff.shift) if in_aff else None
            # d phi for every (target, slot), then back to edge-list order
            dphi_in = torch.empty((g.n_tgt * 4, pk.f_in), dtype=torch.float32, device=dev)
            call("dgnn_upd_edge_bwd", ptr(x_in), sc_in, sh_in, int(relu_in), ptr(g.nbr), ptr(d_agg), None, None, None,
                 g.n_tgt, pk.f_in, ptr(dphi_in), st)
            dphi_e = torch.zeros((n_e, pk.f_in), dtype=torch.float32, device=dev)
            call("dgnn_scatter_rows", ptr(dphi_in), ptr(g._eid_in), g.n_tgt * 4, pk.f_in, ptr(dphi_e), st)
            noaff, nocoef = Affine(None, None, None, None), (None, None, None)
            _, d_h1, db3, dw3 = _dense_and_dw(dphi_e, phi_in, nocoef, noaff, w3p, None, None, e1, aff_e, True, n_e, h1,
                                              pk.f_in)
            grads["convs.%d.e3_w" % l] = dw3[:c.f_in]
            grads["convs.%d.e3_b" % l] = db3[:c.f_in]
            part = torch.empty((small, 2 * h1), dtype=torch.float64, device=dev)
            dy1 = torch.empty_like(d_h1)
            call("dgnn_act_bwd", ptr(d_h1), ptr(e1), ptr(aff_e.scale), ptr(aff_e.shift), ptr(aff_e.mean), ptr(aff_e.rstd),
                 1, n_e, h1, ptr(dy1), ptr(part), st)
            r = _reduce(part)
            s1e, s2e = r[:h1], r[h1:]
            grads["convs.%d.e_norm_w" % l], grads["convs.%d.e_norm_b" % l] = s2e, s1e
            coeffs_e = _norm_coeffs(em.norm, aff_e, s1e, s2e, n_e, h1)
            _, _, db0, dw0 = _dense_and_dw(dy1, e1, coeffs_e, aff_e, w0p, None, None, g.ea_edges, None, False, n_e,
                                           g.ea_edges.shape[1], h1)
            grads["convs.%d.e0_w" % l] = dw0[:, :em.w0.shape[1]]
            grads["convs.%d.e0_b" % l] = db0
            if need_prev:
                dh = torch.empty((g.n_src, pk.f_in), dtype=torch.float32, device=dev)
                call("dgnn_gather_phi_bwd", ptr(d_agg), ptr(d_self), ptr(g.onbr), ptr(g.orow()), ptr(phi_in), None, 0,
                     g.n_src, g.n_tgt, pk.f_in, ptr(dh), st)
                part = torch.empty((small, 2 * pk.f_in), dtype=torch.float64, device=dev)
                call("dgnn_act_bwd", ptr(dh), ptr(x_in), sc_in, sh_in, ptr(in_aff.mean) if in_aff else None,
                     ptr(in_aff.rstd) if in_aff else None, int(relu_in), g.n_src, pk.f_in, ptr(dh), ptr(part), st)
                r = _reduce(part)
                s1, s2 = r[:pk.f_in], r[pk.f_in:]
                dy = dh
        elif tc_gather:
            # dh / dy_prev / (S1,S2) and dW_e / db_e with the edge filter on tensor cores
            tcg = lib().dgnn_tc_grid()
            part = torch.empty((tcg, 2 * pk.f_in), dtype=torch.float64, device=dev) if need_prev else None
            dy_prev = torch.empty((n_srcs, pk.f_in), dtype=torch.float32, device=dev) if need_prev else None
            dwe_p = torch.empty((tcg, pk.f_in, 32), dtype=torch.float32, device=dev)
            call("dgnn_gather_tc_bwd", ptr(d_agg), ptr(d_self), ptr(g.onbr), ptr(g.ea_own), pk.fe, ptr(pk.w_e),
                 ptr(pk.b_e), ptr(x_in), ptr(in_aff.scale) if in_aff else None, ptr(in_aff.shift) if in_aff else None,
                 ptr(in_aff.mean) if in_aff else None, ptr(in_aff.rstd) if in_aff else None, int(relu_in),
                 n_srcs, g.n_tgt, pk.f_in, ptr(dy_prev), ptr(part), ptr(dwe_p), st)
            dwe = torch.empty((pk.f_in, 32), dtype=torch.float32, device=dev)
            call("dgnn_reduce_partials_f32", ptr(dwe_p), tcg, pk.f_in * 32, ptr(dwe), st)
            fe_u = c.w_e.shape[1]
            grads["convs.%d.w_e" % l] = dwe[:c.f_in, :fe_u]
            grads["convs.%d.b_e" % l] = dwe[:c.f_in, pk.fe]
            if need_prev:
                r = _reduce(part)
                s1, s2 = r[:pk.f_in], r[pk.f_in:]
                dy = dy_prev
        elif need_prev or pk.fe:
            grid = lib().dgnn_gather_bwd_grid(pk.f_in)
            plen = pk.f_in * (pk.fe + 1) + 2 * pk.f_in
            part = torch.empty((grid, plen), dtype=torch.float64, device=dev)
            dy_prev = torch.empty((n_srcs, pk.f_in), dtype=torch.float32, device=dev) if need_prev else None
            call("dgnn_gather_bwd", ptr(d_agg), ptr(d_self), ptr(g.onbr), ptr(g.ea_own) if pk.fe else None, pk.fe,
                 ptr(pk.w_e), ptr(pk.b_e), ptr(x_in), ptr(in_aff.scale) if in_aff else None,
                 ptr(in_aff.shift) if in_aff else None, ptr(in_aff.mean) if in_aff else None,
                 ptr(in_aff.rstd) if in_aff else None, int(relu_in), n_srcs, g.n_tgt, pk.f_in, ptr(dy_prev),
                 ptr(part), st)
            r = _reduce(part)
            if pk.fe:
                fe_u = c.w_e.shape[1]
                grads["convs.%d.w_e" % l] = r[:pk.f_in * pk.fe].view(pk.f_in, pk.fe)[:c.f_in, :fe_u]
                grads["convs.%d.b_e" % l] = r[pk.f_in * pk.fe:pk.f_in * (pk.fe + 1)][:c.f_in]
            if need_prev:
                s1 = r[pk.f_in * (pk.fe + 1):pk.f_in * (pk.fe + 1) + pk.f_in]
                s2 = r[pk.f_in * (pk.fe + 1) + pk.f_in:]
                dy = dy_prev
    return grads
