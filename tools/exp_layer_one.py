"""Dev: every hot kernel of one layer (forward + backward), a few launches each, for ncu / timing.
    python tools/exp_layer_one.py [n_objects] [reps] [f_in] [f_out]        (default 128 -> 128; kf96 also has 28 -> 64, 64 -> 128)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgnn_b200._lib import call, ptr, lib
from dgnn_b200.graph import build_full_graph
from dgnn_b200 import engine
import bench
DEV = "cuda:0"
nobj = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
host = bench.make_objects(nobj, 0)
n = host["n"]
eg = build_full_graph(host["edge_index"], host["edge_attr"], n, DEV, pos=host["pos"], order="morton")
f = int(sys.argv[3]) if len(sys.argv) > 3 else 128          # input width (multiple of 4)
fo = int(sys.argv[4]) if len(sys.argv) > 4 else 128         # output width
fe = 20
x = torch.randn(n, f, device=DEV); w_e = torch.randn(f, fe, device=DEV) * 0.3; b_e = torch.randn(f, device=DEV)
sc = torch.rand(f, device=DEV) + 0.5; sh = torch.randn(f, device=DEV) * 0.1
mean = torch.randn(f, device=DEV) * 0.1; rstd = torch.rand(f, device=DEV) + 0.5
mean_o = torch.randn(fo, device=DEV) * 0.1; rstd_o = torch.rand(fo, device=DEV) + 0.5
g3 = torch.rand(3, fo, device=DEV)
agg = torch.empty(n, f, device=DEV); z = torch.empty(n, fo, device=DEV); dy = torch.randn(n, fo, device=DEV)
d_agg = torch.empty(n, f, device=DEV); d_self = torch.empty(n, f, device=DEV); dyp = torch.empty(n, f, device=DEV)
w_cat = torch.randn(fo, 2 * f, device=DEV) * 0.05; bias = torch.randn(fo, device=DEV)
st = torch.cuda.current_stream().cuda_stream
tcg = lib().dgnn_tc_grid()
b_fwd = engine.pack_b(w_cat, fo, f, 2)
b_bwd = engine.pack_b(w_cat.t().contiguous(), 2 * f, fo, 1, backward=True)
stats = torch.empty((tcg, 2, fo), dtype=torch.float64, device=DEV)
db_p = torch.empty((tcg, fo), dtype=torch.float64, device=DEV)
dw_p = torch.empty((tcg, fo, 2 * f), device=DEV)
part = torch.empty((tcg, 2 * f), dtype=torch.float64, device=DEV)
dwe = torch.empty((tcg, f, 32), device=DEV)
steps = [
    ("gather_fwd", lambda: call("dgnn_gather_tc_fwd", ptr(x), ptr(sc), ptr(sh), 1, ptr(eg.nbr), ptr(eg.ea_in), fe, ptr(w_e), ptr(b_e), n, f, ptr(agg), st)),
    ("dense_fwd", lambda: call("dgnn_dense_fwd_tc", ptr(agg), ptr(x), ptr(sc), ptr(sh), 1, ptr(b_fwd), ptr(bias), None, None, 0, n, f, fo, ptr(z), ptr(stats), st)),
    ("dense_bwd", lambda: call("dgnn_dense_bwd_tc", ptr(dy), ptr(z), ptr(g3[0]), ptr(g3[1]), ptr(g3[2]), ptr(mean_o), ptr(rstd_o), ptr(b_bwd), ptr(eg.nbr), n, f, fo, ptr(d_agg), ptr(d_self), None, st)),
    ("dw_bwd", lambda: call("dgnn_dw_bwd_tc", ptr(dy), ptr(z), ptr(g3[0]), ptr(g3[1]), ptr(g3[2]), ptr(mean_o), ptr(rstd_o), ptr(agg), ptr(x), ptr(sc), ptr(sh), 1, n, f, fo, 2 * f, ptr(dw_p), ptr(db_p), st)),
    ("gather_bwd", lambda: call("dgnn_gather_tc_bwd", ptr(d_agg), ptr(d_self), ptr(eg.onbr), ptr(eg.ea_own), fe, ptr(w_e), ptr(b_e), ptr(x), ptr(sc), ptr(sh), ptr(mean), ptr(rstd), 1, n, n, f, ptr(dyp), ptr(part), None, st)),
    ("dwe_bwd", lambda: call("dgnn_gather_tc_bwd", ptr(d_agg), ptr(d_self), ptr(eg.onbr), ptr(eg.ea_own), fe, ptr(w_e), ptr(b_e), ptr(x), ptr(sc), ptr(sh), ptr(mean), ptr(rstd), 1, n, n, f, None, None, ptr(dwe), st)),
]
print("cells", n, "layer %d -> %d" % (f, fo))
for name, fn in steps:
    for _ in range(0 if os.environ.get('NCU') else 2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-12s %.3f ms  %.2f us / 128-cell tile / SM" % (name, ms, ms * 1e3 / (n / 128 / 148)))
