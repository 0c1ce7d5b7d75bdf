"""Dev: a few launches of the forward / backward gather kernels at F=128 for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgnn_b200._lib import call, ptr, lib
from dgnn_b200.graph import build_full_graph
import bench
DEV = "cuda:0"
host = bench.make_objects(32, 0)
n = host["n"]
eg = build_full_graph(host["edge_index"], host["edge_attr"], n, DEV, pos=host["pos"], order="morton")
f, fe = 128, 20
x = torch.randn(n, f, device=DEV); w_e = torch.randn(f, fe, device=DEV) * 0.3; b_e = torch.randn(f, device=DEV)
sc = torch.rand(f, device=DEV) + 0.5; sh = torch.randn(f, device=DEV) * 0.1
agg = torch.empty(n, f, device=DEV); d_self = torch.randn(n, f, device=DEV)
st = torch.cuda.current_stream().cuda_stream
tcg = lib().dgnn_tc_grid()
part = torch.empty((tcg, 2 * f), dtype=torch.float64, device=DEV)
dwe = torch.empty((tcg, f, 32), device=DEV)
mean = torch.randn(f, device=DEV) * 0.1; rstd = torch.rand(f, device=DEV) + 0.5
for _ in range(3):
    call("dgnn_gather_tc_fwd", ptr(x), ptr(sc), ptr(sh), 1, ptr(eg.nbr), ptr(eg.ea_in), fe, ptr(w_e), ptr(b_e), n, f, ptr(agg), st)
    call("dgnn_gather_tc_bwd", ptr(x), ptr(d_self), ptr(eg.onbr), ptr(eg.ea_own), fe, ptr(w_e), ptr(b_e), ptr(agg), ptr(sc), ptr(sh),
         ptr(mean), ptr(rstd), 1, n, n, f, ptr(agg), ptr(part), ptr(dwe), st)
torch.cuda.synchronize()
print("ok")
