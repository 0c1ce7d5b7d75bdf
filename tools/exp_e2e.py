"""Dev experiment: breakdown of the end-to-end (host buffers) training step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dgnn_b200 import runModel as rm
from dgnn_b200.synthetic import make_clf, to_attr
from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
from dgnn_b200.graph import build_full_graph
dev = "cuda:0"
host = bench.make_objects(64, 0)
pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
n = host["n"]
def T(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("cells", n)
print("H2D x          %.2f ms" % T(lambda: pinned["x"].to(dev, non_blocking=True)))
print("H2D edge_attr  %.2f ms" % T(lambda: pinned["edge_attr"].to(dev, non_blocking=True)))
print("H2D edge_index %.2f ms" % T(lambda: pinned["edge_index"].to(dev, non_blocking=True)))
ar = torch.arange(4 * n)
print("host arange(E) %.2f ms" % T(lambda: torch.arange(4 * n)))
print("H2D pageable e_id %.2f ms" % T(lambda: ar.to(dev)))
ei = pinned["edge_index"].to(dev); ea = pinned["edge_attr"].to(dev); pos = pinned["pos"].to(dev)
print("build_full_graph (device inputs, morton) %.2f ms" % T(lambda: build_full_graph(ei, ea, n, dev, pos=pos, order="morton")))
eid = ar.to(dev)
print("ea[e_id] device gather %.2f ms" % T(lambda: ea[eid]))
