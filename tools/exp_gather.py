"""Dev experiment: where does the forward gather kernel spend its time? (flags in relu_in bits)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dgnn_b200 import synthetic as syn
from dgnn_b200._lib import call, ptr
from dgnn_b200.graph import build_full_graph
DEV = "cuda:0"
import bench
host = bench.make_objects(32, 0)
n = host["n"]
eg = build_full_graph(host["edge_index"], host["edge_attr"], n, DEV, pos=host["pos"], order="morton")
f, fe = 128, 20
x = torch.randn(n, f, device=DEV); w_e = torch.randn(f, fe, device=DEV) * 0.3; b_e = torch.randn(f, device=DEV)
sc = torch.rand(f, device=DEV) + 0.5; sh = torch.randn(f, device=DEV) * 0.1
agg = torch.empty(n, f, device=DEV)
st = torch.cuda.current_stream().cuda_stream
def run(flags, nbr=None):
    nb = eg.nbr if nbr is None else nbr
    for _ in range(3):
        call("dgnn_gather_tc_fwd", ptr(x), ptr(sc), ptr(sh), flags, ptr(nb), ptr(eg.ea_in), fe, ptr(w_e), ptr(b_e), n, f, ptr(agg), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call("dgnn_gather_tc_fwd", ptr(x), ptr(sc), ptr(sh), flags, ptr(nb), ptr(eg.ea_in), fe, ptr(w_e), ptr(b_e), n, f, ptr(agg), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
print("cells", n)
for name, fl in (("full", 1), ("no x loads", 3), ("no tmem ld", 5), ("no stores", 9), ("no x, no tmem", 7), ("nothing (x,tmem,store off)", 15)):
    print("%-28s %.3f ms" % (name, run(fl)))
# locality: identity neighbours (each cell gathers itself and next rows) vs real
ident = torch.stack([torch.arange(n, device=DEV, dtype=torch.int32)] * 4, 1).contiguous()
ident[:, 1] = (ident[:, 1] + 1) % n; ident[:, 2] = (ident[:, 2] + 2) % n; ident[:, 3] = (ident[:, 3] + 3) % n
print("%-28s %.3f ms" % ("full, sequential neighbours", run(1, ident)))
rnd = torch.randint(0, n, (n, 4), device=DEV, dtype=torch.int32)
print("%-28s %.3f ms" % ("full, random neighbours", run(1, rnd)))
