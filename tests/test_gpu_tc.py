"""GPU: the tcgen05 (3xTF32) kernels against float64 references and against the generic FP32
kernels, called through the C ABI."""
import numpy as np
import pytest
import torch

from tests.helpers import make_graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("n,f_in,f_out", [(1000, 28, 64), (129, 64, 128), (4096, 128, 128), (777, 128, 64), (50, 32, 32),
                                          # tensor-map edge cases: one row, rows / columns beyond the boxes, several
                                          # output slices, a ring longer than a tile
                                          (1, 4, 4), (127, 36, 132), (128, 100, 260), (31, 256, 40), (20000, 128, 128)])
def test_dense_forward_tc_matches_fp64(n, f_in, f_out):
    from dgnn_b200 import engine
    torch.manual_seed(0)
    x = torch.randn(n, f_in, device=DEV)
    w = torch.randn(f_out, f_in, device=DEV) * 0.2
    b = torch.randn(f_out, device=DEV)
    sc = torch.rand(f_in, device=DEV) + 0.5
    sh = torch.randn(f_in, device=DEV) * 0.3
    aff = engine.Affine(sc, sh)
    bp = engine.pack_b(w, f_out, f_in, 1)
    out, _, stats = engine._layer_fwd(x, aff, True, None, None, b, None, None, 0, None, False, n, f_in, f_out,
                                      False, True, b_packed=bp)
    h = torch.relu(x.double() * sc.double() + sh.double())
    ref = h @ w.double().t() + b.double()
    assert _rel(out, ref) < 2e-6
    # per-channel (sum, sum^2) partials
    st = stats.sum(0)
    # a column sum cancels: its error is measured against the sum of magnitudes (fp32 per-tile partials, doubles after)
    np.testing.assert_allclose(st[0].cpu().numpy(), ref.sum(0).cpu().numpy(), rtol=1e-5,
                               atol=max(1e-3, 2e-6 * ref.abs().sum(0).max().item()))
    np.testing.assert_allclose(st[1].cpu().numpy(), (ref * ref).sum(0).cpu().numpy(), rtol=1e-5)
    # eval epilogue: affine + relu fused
    osc = torch.rand(f_out, device=DEV) + 0.5
    osh = torch.randn(f_out, device=DEV)
    out2, _, _ = engine._layer_fwd(x, aff, True, None, None, b, None, None, 0, engine.Affine(osc, osh), True, n, f_in,
                                   f_out, False, False, b_packed=bp)
    ref2 = torch.relu(ref * osc.double() + osh.double())
    assert _rel(out2, ref2) < 2e-6


@pytest.mark.parametrize("f_in,f_out,fe", [(28, 64, 20), (64, 128, 20), (128, 128, 20), (32, 32, 0)])
def test_gather_forward_tc_matches_generic_kernel(f_in, f_out, fe):
    from dgnn_b200 import engine
    from dgnn_b200.graph import build_full_graph
    g = make_graph(900, seed=3)
    n = g["n"]
    ei = torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()
    torch.manual_seed(1)
    ea = torch.randn(4 * n, fe) if fe else None
    eg = build_full_graph(ei, ea, n, DEV, order="rcm")
    x = torch.randn(n, f_in, device=DEV)
    w_cat = torch.randn(f_out, 2 * f_in, device=DEV) * 0.2
    bias = torch.randn(f_out, device=DEV)
    w_e = torch.randn(f_in, fe, device=DEV) * 0.3 if fe else None
    b_e = torch.randn(f_in, device=DEV) if fe else None
    sc = torch.rand(f_in, device=DEV) + 0.5
    sh = torch.randn(f_in, device=DEV) * 0.3
    aff = engine.Affine(sc, sh)
    wt = w_cat.t().contiguous()
    o1, a1, s1 = engine._layer_fwd(x, aff, True, eg, wt, bias, w_e, b_e, fe, None, False, n, f_in, f_out, True, True)
    bp = engine.pack_b(w_cat, f_out, f_in, 2)
    o2, a2, s2 = engine._layer_fwd(x, aff, True, eg, wt, bias, w_e, b_e, fe, None, False, n, f_in, f_out, True, True,
                                   b_packed=bp)
    assert _rel(a2, a1) < 1e-6
    assert _rel(o2, o1) < 5e-6
    np.testing.assert_allclose(s2.sum(0).cpu().numpy(), s1.sum(0).cpu().numpy(), rtol=2e-5, atol=1e-2)


@pytest.mark.parametrize("with_db", [True, False], ids=["db", "nodb"])   # nodb: the TMA tensor-map kernel (db from dW)
@pytest.mark.parametrize("n,f_in,f_out,gather", [(1000, 28, 64, True), (3000, 128, 128, True), (515, 128, 64, False),
                                                  (640, 64, 128, True), (1, 32, 4, True), (127, 64, 36, True),
                                                  (129, 100, 132, False), (130, 160, 300, True), (20000, 128, 128, True)])
def test_dense_backward_tc_matches_fp64(n, f_in, f_out, gather, with_db):
    from dgnn_b200 import engine
    from dgnn_b200._lib import call, lib, ptr
    torch.manual_seed(2)
    dy = torch.randn(n, f_out, device=DEV)
    z = torch.randn(n, f_out, device=DEV)
    gq, aq, bq = (torch.randn(f_out, device=DEV) for _ in range(3))
    mean = torch.randn(f_out, device=DEV) * 0.1
    rstd = torch.rand(f_out, device=DEV) + 0.5
    k_total = 2 * f_in if gather else f_in
    w_cat = torch.randn(f_out, k_total, device=DEV) * 0.2
    nbr = None
    if gather:
        nbr = torch.randint(0, n, (n, 4), device=DEV, dtype=torch.int32)
        nbr[::5, 3] = -1
        nbr[::7, 1] = -1
    bp = engine.pack_b(w_cat.t().contiguous(), k_total, f_out, 1, backward=True)
    d_self = torch.empty(n, f_in, device=DEV)
    d_agg = torch.empty(n, f_in, device=DEV) if gather else None
    if with_db and f_out > 256:
        pytest.skip("the shared column sums cover f_out <= 256")
    db_p = torch.empty(lib().dgnn_tc_grid(), f_out, dtype=torch.float64, device=DEV) if with_db else None
    st = torch.cuda.current_stream().cuda_stream
    call("dgnn_dense_bwd_tc", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(mean), ptr(rstd), ptr(bp), ptr(nbr), n,
         f_in, f_out, ptr(d_agg), ptr(d_self), ptr(db_p), st)
    dz = gq.double() * dy.double() - (aq.double() + (z.double() - mean.double()) * rstd.double() * bq.double())
    dA = dz @ w_cat.double()
    if gather:
        cnt = (nbr >= 0).sum(1).clamp(min=1).double()
        assert _rel(d_agg, dA[:, :f_in] / cnt[:, None]) < 3e-6
        assert _rel(d_self, dA[:, f_in:]) < 3e-6
    else:
        assert _rel(d_self, dA) < 3e-6
    if with_db:
        np.testing.assert_allclose(db_p.sum(0).cpu().numpy(), dz.sum(0).cpu().numpy(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("n,f_in,f_out,gather", [(1000, 28, 64, True), (5000, 128, 128, True), (515, 128, 64, False),
                                                  (31, 64, 128, True), (40000, 64, 128, True), (1, 4, 4, False),
                                                  (33, 28, 64, True), (100, 128, 64, False), (30000, 36, 100, True)])
def test_dw_tc_matches_fp64(n, f_in, f_out, gather):
    from dgnn_b200._lib import call, lib, ptr
    torch.manual_seed(3)
    dy = torch.randn(n, f_out, device=DEV)
    z = torch.randn(n, f_out, device=DEV)
    gq, aq, bq = (torch.randn(f_out, device=DEV) for _ in range(3))
    mean = torch.randn(f_out, device=DEV) * 0.1
    rstd = torch.rand(f_out, device=DEV) + 0.5
    x = torch.randn(n, f_in, device=DEV)
    sc = torch.rand(f_in, device=DEV) + 0.5
    sh = torch.randn(f_in, device=DEV) * 0.3
    agg = torch.randn(n, f_in, device=DEV) if gather else None
    k_total = 2 * f_in if gather else f_in
    assert lib().dgnn_dw_tc_supported(f_out, k_total)
    part = torch.empty(lib().dgnn_tc_grid(), f_out, k_total, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    db_p = torch.empty(lib().dgnn_tc_grid(), f_out, dtype=torch.float64, device=DEV)
    call("dgnn_dw_bwd_tc", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(mean), ptr(rstd), ptr(agg), ptr(x), ptr(sc),
         ptr(sh), 1, n, f_in, f_out, k_total, ptr(part), ptr(db_p), st)
    dw = part.double().sum(0)
    dz = gq.double() * dy.double() - (aq.double() + (z.double() - mean.double()) * rstd.double() * bq.double())
    h = torch.relu(x.double() * sc.double() + sh.double())
    A = torch.cat([agg.double(), h], 1) if gather else h
    ref = dz.t() @ A
    assert _rel(dw, ref) < 5e-6
    np.testing.assert_allclose(db_p.sum(0).cpu().numpy(), dz.sum(0).cpu().numpy(), rtol=1e-4, atol=1e-3)   # bias gradient


@pytest.mark.parametrize("f_in,fe,ragged", [(28, 20, False), (64, 20, True), (128, 20, False), (128, 4, True)])
def test_gather_tc_forward_and_backward_match_generic_kernels(f_in, fe, ragged):
    from dgnn_b200 import engine
    from dgnn_b200._lib import call, lib, ptr
    from dgnn_b200.graph import build_full_graph
    g = make_graph(1100, seed=4)
    n = g["n"]
    ei = torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()
    torch.manual_seed(5)
    ea = torch.randn(4 * n, fe)
    eg = build_full_graph(ei, ea, n, DEV, order="rcm")
    if ragged:
        eg.nbr = eg.nbr.clone(); eg.onbr = eg.nbr
        eg.nbr[::3, 2] = -1
        eg.nbr[::11] = -1
    x = torch.randn(n, f_in, device=DEV)
    w_e = torch.randn(f_in, fe, device=DEV) * 0.3
    b_e = torch.randn(f_in, device=DEV)
    sc = torch.rand(f_in, device=DEV) + 0.5
    sh = torch.randn(f_in, device=DEV) * 0.3
    st = torch.cuda.current_stream().cuda_stream
    # forward: agg
    f_out = 32
    wt = torch.zeros(2 * f_in, f_out, device=DEV)
    _, agg_ref, _ = engine._layer_fwd(x, engine.Affine(sc, sh), True, eg, wt, None, w_e, b_e, fe, None, False, n, f_in,
                                      f_out, True, False)
    agg = torch.empty(n, f_in, device=DEV)
    call("dgnn_gather_tc_fwd", ptr(x), ptr(sc), ptr(sh), 1, ptr(eg.nbr), ptr(eg.ea_in), fe, ptr(w_e), ptr(b_e), n, f_in,
         ptr(agg), st)
    assert _rel(agg, agg_ref) < 3e-6
    # backward: dy_prev, S1, S2
    d_agg = torch.randn(n, f_in, device=DEV)
    d_self = torch.randn(n, f_in, device=DEV)
    mean = torch.randn(f_in, device=DEV) * 0.1
    rstd = torch.rand(f_in, device=DEV) + 0.5
    grid = lib().dgnn_gather_bwd_grid(f_in)
    plen = f_in * (fe + 1) + 2 * f_in
    part = torch.empty(grid, plen, dtype=torch.float64, device=DEV)
    dy_ref = torch.empty(n, f_in, device=DEV)
    call("dgnn_gather_bwd", ptr(d_agg), ptr(d_self), ptr(eg.onbr), ptr(eg.ea_own), fe, ptr(w_e), ptr(b_e), ptr(x), ptr(sc),
         ptr(sh), ptr(mean), ptr(rstd), 1, n, n, f_in, ptr(dy_ref), ptr(part), st)
    r = part.sum(0)
    s1_ref, s2_ref = r[f_in * (fe + 1):f_in * (fe + 1) + f_in], r[f_in * (fe + 1) + f_in:]
    dwe_ref = r[:f_in * (fe + 1)]
    dy = torch.empty(n, f_in, device=DEV)
    tcg = lib().dgnn_tc_grid()
    part2 = torch.empty(tcg, 2 * f_in, dtype=torch.float64, device=DEV)
    dwe_p = torch.empty(tcg, f_in, 32, device=DEV)
    call("dgnn_gather_tc_bwd", ptr(d_agg), ptr(d_self), ptr(eg.onbr), ptr(eg.ea_own), fe, ptr(w_e), ptr(b_e), ptr(x),
         ptr(sc), ptr(sh), ptr(mean), ptr(rstd), 1, n, n, f_in, ptr(dy), ptr(part2), ptr(dwe_p), st)
    assert _rel(dy, dy_ref) < 3e-6
    r2 = part2.sum(0)
    assert _rel(r2[:f_in], s1_ref) < 1e-5 and _rel(r2[f_in:], s2_ref) < 1e-5
    # edge-filter gradients: dW_e (dphi rounded to TF32 -> ~1e-4 relative) and db_e (column fe)
    dwe = dwe_p.double().sum(0)
    assert _rel(dwe[:, :fe], dwe_ref[:f_in * fe].view(f_in, fe)) < 3e-4
    assert _rel(dwe[:, fe], dwe_ref[f_in * fe:]) < 3e-4
    # FP32 edge-filter-only kernel
    part3 = torch.empty(grid, plen, dtype=torch.float64, device=DEV)
    call("dgnn_edge_filter_bwd", ptr(d_agg), ptr(eg.onbr), ptr(eg.ea_own), fe, ptr(w_e), ptr(b_e), ptr(x), ptr(sc), ptr(sh),
         1, n, n, f_in, ptr(part3), st)
    assert _rel(part3.sum(0)[:f_in * (fe + 1)], dwe_ref) < 1e-6


def test_gather_tc_kernels_are_bitwise_reproducible():
    """The elect-by-arrival MMA issue must not leak timing into the results: forward aggregate, backward dy / (S1, S2)
    partials and the dW_e partials of repeated launches on a graph that spans many tiles are bit-identical."""
    from dgnn_b200._lib import call, lib, ptr
    from dgnn_b200.graph import build_full_graph
    g = make_graph(9000, seed=9)
    n, f_in, fe = g["n"], 128, 20
    ei = torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()
    torch.manual_seed(6)
    eg = build_full_graph(ei, torch.randn(4 * n, fe), n, DEV, order="rcm")
    x = torch.randn(n, f_in, device=DEV)
    w_e = torch.randn(f_in, fe, device=DEV) * 0.3
    b_e = torch.randn(f_in, device=DEV)
    sc = torch.rand(f_in, device=DEV) + 0.5
    sh = torch.randn(f_in, device=DEV) * 0.3
    mean = torch.randn(f_in, device=DEV) * 0.1
    rstd = torch.rand(f_in, device=DEV) + 0.5
    d_agg = torch.randn(n, f_in, device=DEV)
    d_self = torch.randn(n, f_in, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    tcg = lib().dgnn_tc_grid()

    def once():
        agg = torch.empty(n, f_in, device=DEV)
        dy = torch.empty(n, f_in, device=DEV)
        part = torch.empty(tcg, 2 * f_in, dtype=torch.float64, device=DEV)
        dwe = torch.empty(tcg, f_in, 32, device=DEV)
        call("dgnn_gather_tc_fwd", ptr(x), ptr(sc), ptr(sh), 1, ptr(eg.nbr), ptr(eg.ea_in), fe, ptr(w_e), ptr(b_e), n, f_in,
             ptr(agg), st)
        call("dgnn_gather_tc_bwd", ptr(d_agg), ptr(d_self), ptr(eg.onbr), ptr(eg.ea_own), fe, ptr(w_e), ptr(b_e), ptr(x),
             ptr(sc), ptr(sh), ptr(mean), ptr(rstd), 1, n, n, f_in, ptr(dy), ptr(part), ptr(dwe), st)
        torch.cuda.synchronize()
        return agg, dy, part, dwe

    ref = once()
    for _ in range(8):
        for a, b in zip(once(), ref):
            assert torch.equal(a, b)


def test_dense_and_dw_tc_kernels_are_bitwise_reproducible():
    """One MMA-issuing thread per CTA accumulates every tile in a fixed order (several issuing warps were measured and
    dropped for exactly this reason): z, the BatchNorm partials, d_agg / d_self and the dW / db
    partials of repeated launches are bit-identical."""
    from dgnn_b200 import engine
    from dgnn_b200._lib import call, lib, ptr
    torch.manual_seed(8)
    n, f_in, f_out = 40000, 128, 128
    x = torch.randn(n, f_in, device=DEV); agg = torch.randn(n, f_in, device=DEV)
    sc = torch.rand(f_in, device=DEV) + 0.5; sh = torch.randn(f_in, device=DEV) * 0.3
    w_cat = torch.randn(f_out, 2 * f_in, device=DEV) * 0.1; bias = torch.randn(f_out, device=DEV)
    dy = torch.randn(n, f_out, device=DEV); z = torch.randn(n, f_out, device=DEV)
    gq, aq, bq = (torch.randn(f_out, device=DEV) for _ in range(3))
    mean = torch.randn(f_out, device=DEV) * 0.1; rstd = torch.rand(f_out, device=DEV) + 0.5
    nbr = torch.randint(0, n, (n, 4), device=DEV, dtype=torch.int32)
    b_fwd = engine.pack_b(w_cat, f_out, f_in, 2)
    b_bwd = engine.pack_b(w_cat.t().contiguous(), 2 * f_in, f_out, 1, backward=True)
    st = torch.cuda.current_stream().cuda_stream
    tcg = lib().dgnn_tc_grid()

    def once():
        out = torch.empty(n, f_out, device=DEV)
        stats = torch.empty(tcg, 2, f_out, dtype=torch.float64, device=DEV)
        d_agg = torch.empty(n, f_in, device=DEV); d_self = torch.empty(n, f_in, device=DEV)
        dw = torch.empty(tcg, f_out, 2 * f_in, device=DEV)
        db = torch.empty(tcg, f_out, dtype=torch.float64, device=DEV)
        call("dgnn_dense_fwd_tc", ptr(agg), ptr(x), ptr(sc), ptr(sh), 1, ptr(b_fwd), ptr(bias), None, None, 0, n, f_in, f_out,
             ptr(out), ptr(stats), st)
        call("dgnn_dense_bwd_tc", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(mean), ptr(rstd), ptr(b_bwd), ptr(nbr), n,
             f_in, f_out, ptr(d_agg), ptr(d_self), None, st)
        call("dgnn_dw_bwd_tc", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(mean), ptr(rstd), ptr(agg), ptr(x), ptr(sc),
             ptr(sh), 1, n, f_in, f_out, 2 * f_in, ptr(dw), ptr(db), st)
        torch.cuda.synchronize()
        return out, stats, d_agg, d_self, dw, db

    ref = once()
    for _ in range(6):
        for a, b in zip(once(), ref):
            assert torch.equal(a, b)
