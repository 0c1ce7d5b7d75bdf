import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgnn_b200 import engine
from dgnn_b200._lib import call, lib, ptr
DEV = "cuda:0"
torch.manual_seed(2)
for (n, f_in, f_out) in ((1000, 28, 64), (2498, 128, 128), (2498, 128, 64)):
    dy = torch.randn(n, f_out, device=DEV); z = torch.randn(n, f_out, device=DEV)
    gq, aq, bq = (torch.randn(f_out, device=DEV) for _ in range(3))
    mean = torch.randn(f_out, device=DEV) * 0.1; rstd = torch.rand(f_out, device=DEV) + 0.5
    k_total = 2 * f_in
    w_cat = torch.randn(f_out, k_total, device=DEV) * 0.2
    nbr = torch.randint(0, n, (n, 4), device=DEV, dtype=torch.int32)
    bp = engine.pack_b(w_cat.t().contiguous(), k_total, f_out, 1, backward=True)
    d_self = torch.empty(n, f_in, device=DEV); d_agg = torch.empty(n, f_in, device=DEV)
    db_p = torch.empty(lib().dgnn_tc_grid(), f_out, dtype=torch.float64, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    call("dgnn_dense_bwd_tc", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(mean), ptr(rstd), ptr(bp), ptr(nbr), n, f_in, f_out, ptr(d_agg), ptr(d_self), ptr(db_p), st)
    dz = gq.double() * dy.double() - (aq.double() + (z.double() - mean.double()) * rstd.double() * bq.double())
    ref = dz.sum(0); got = db_p.sum(0)
    err = (got - ref).abs()
    print(n, f_in, f_out, "max db err", err.max().item(), "sum|dz| col max", dz.abs().sum(0).max().item(), "argmax col", err.argmax().item())
    print("  first cols err:", [float("%.2e" % e) for e in err[:8].tolist()], " per-cta nonzero rows:", int((db_p.abs().sum(1) > 0).sum()))
    db2 = torch.empty(lib().dgnn_layer_grid(f_in, f_out), f_out, dtype=torch.float64, device=DEV)
    call("dgnn_dense_bwd", ptr(dy), ptr(z), ptr(gq), ptr(aq), ptr(bq), ptr(mean), ptr(rstd), ptr(w_cat), ptr(nbr), n, f_in, f_out, k_total, ptr(d_agg), ptr(d_self), ptr(db2), st)
    print("  fma db err", (db2.sum(0) - ref).abs().max().item())
