#!/usr/bin/env python
"""Benchmark of the DGNN cell-classification hot path on B200 (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): ModelNet10-shaped
training — a batch is the disjoint union of B synthetic object graphs (scipy Delaunay of 3 000
scan-like points each, ~19.7 k cells per object incl. infinite cells, random features of the
``feat`` tool's shape), kf96 widths 28->64->128->128->128, decoder 128->64->2, BatchNorm in
training mode, volume-weighted KL loss, Adam lr 0.005.  One step = forward + loss + backward +
Adam over the whole batch.  Metric: Delaunay cells/s (whole job, all GPUs).

N > 1 (launched by torchrun, one rank per GPU): ONE scene graph partitioned over the GPUs (north_star
subsystem 4, BASELINE configs[3]/[4]): every rank generates its own shard of an analytic 4-regular
lattice scene on the device (~8.4 M cells per GPU: 16.8 M / 33.5 M / 67.1 M cells at 2 / 4 / 8 GPUs), the
halo maps are negotiated between the ranks (no rank holds the whole graph), owned cells are ordered
boundary-first and every layer's halo exchange (NCCL all-to-all over NVLink) overlaps the interior rows.
`value` = fwd+loss+bwd+Adam cells/s of the Static kf96 model on that scene (the metric's fwd+bwd; batch
statistics, loss normaliser and gradients all-reduced).  Extra keys: `scene_inference` = eval inference on
the FIXED 67.1 M-cell scene at every N (strong scaling; N = 1 carries the single-GPU number of the same
scene), `partition_vs_single_max_abs` = the partitioned logits against the single-GPU path on a small
scene, `dp` = the round-1 data-parallel object-batch number.

``--impl reference`` times the CPU restatement of the reference (oracle/, plain PyTorch with all
host threads; the reference's own modules need torch_geometric, which cannot be installed here) on
a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POINTS_PER_OBJECT = 3000
FE, F0 = 20, 28
WIDTHS = (64, 128, 128, 128)


# --------------------------------------------------------------------------- workload


def make_objects(n_objects: int, seed0: int):
    """Disjoint union of ``n_objects`` synthetic object graphs in the reference's collated layout
    (run.py:59-61): x[N,1+28], edge_attr[4N,20], y[N,2], edge_index int64[2,4N], centroids."""
    from dgnn_b200 import synthetic as og
    xs, eas, ys, eis, cens = [], [], [], [], []
    off = 0
    for i in range(n_objects):
        pts = og.scan_like_points(POINTS_PER_OBJECT, seed=seed0 + i)
        adj, infinite, cen, _ = og.delaunay_graph(pts)
        n = infinite.shape[0]
        x, ea, y = og.synthetic_features(n, infinite, seed=10_000 + seed0 + i)
        xs.append(x); eas.append(ea); ys.append(y); cens.append(cen.astype(np.float32) + 2.0 * i)
        eis.append(adj.T.astype(np.int64) + off)
        off += n
    return dict(x=torch.from_numpy(np.concatenate(xs)), edge_attr=torch.from_numpy(np.concatenate(eas)),
                y=torch.from_numpy(np.concatenate(ys)), edge_index=torch.from_numpy(np.concatenate(eis, axis=1)),
                pos=torch.from_numpy(np.concatenate(cens)), n=off)


def _metric_name():
    """BASELINE.json's metric string (the driver compares it); the fallback is its fwd+bwd part."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "Delaunay cells/sec (GNN fwd+bwd)"


METRIC = _metric_name()


def batch_of(d, to_attr):
    n = d["n"]
    all_ = to_attr({k: v for k, v in d.items() if k != "n"})
    ei = all_.edge_index
    return to_attr(dict(all=all_, batch_n_id=torch.arange(n, device=ei.device),
                        batch_adjs=[(ei, torch.arange(ei.shape[1], device=ei.device), (n, n))] * 5))


# --------------------------------------------------------------------------- clocks


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------- roofline bookkeeping


def algorithmic_bytes(name, a):
    """Compulsory HBM bytes of one launch (DESIGN.md section 'kernels'; SURVEY.md 8d per-layer
    figures split per kernel).  ``a`` = the ctypes argument list of the call."""
    if name in ("dgnn_layer_fwd", "dgnn_layer_fwd_tc"):
        n, f_in, f_out, fe = a[14], a[15], a[16], (a[6] if a[7] else 0)
        gather = a[4] is not None
        b = 4 * f_in + 4 * f_out
        if gather:
            b += 16 + 16 * fe
        if a[18] is not None:
            b += 4 * f_in
        return n * b, "f%d->%d" % (f_in, f_out)
    if name == "dgnn_gather_tc_fwd":
        n, f_in, fe = a[9], a[10], a[6]
        return n * (16 + 16 * fe + 8 * f_in), "f%d" % f_in
    if name == "dgnn_dense_fwd_tc":
        n, f_in, f_out = a[10], a[11], a[12]
        return n * ((8 if a[0] is not None else 4) * f_in + 4 * f_out), "f%d->%d" % (f_in, f_out)
    if name == "dgnn_dense_bwd":
        n, f_in, f_out = a[9], a[10], a[11]
        gather = a[8] is not None
        return n * (8 * f_out + (8 * f_in + 16 if gather else 4 * f_in)), "f%d->%d" % (f_in, f_out)
    if name == "dgnn_dense_bwd_tc":
        n, f_in, f_out = a[9], a[10], a[11]
        gather = a[8] is not None
        return n * (8 * f_out + (8 * f_in + 16 if gather else 4 * f_in)), "f%d->%d" % (f_in, f_out)
    if name in ("dgnn_dw_bwd", "dgnn_dw_bwd_tc"):
        n, f_in, f_out = a[12], a[13], a[14]
        return n * (8 * f_out + (8 * f_in if a[7] is not None else 4 * f_in)), "f%d->%d" % (f_in, f_out)
    if name == "dgnn_gather_bwd":
        n, f_in, fe = a[13], a[15], (a[4] if a[5] else 0)
        return n * (16 + 16 * fe + 12 * f_in + (4 * f_in if a[16] is not None else 0)), "f%d" % f_in
    if name == "dgnn_gather_tc_bwd":
        n, f_in, fe = a[13], a[15], a[4]
        return n * (16 + 16 * fe + 16 * f_in), "f%d" % f_in
    if name == "dgnn_edge_filter_bwd":
        n, f_in, fe = a[10], a[12], a[3]
        return n * (16 + 16 * fe + 8 * f_in), "f%d" % f_in
    return None, ""


_CELLS_ARG = {"dgnn_gather_tc_fwd": 9, "dgnn_dense_fwd_tc": 10, "dgnn_dense_bwd_tc": 9, "dgnn_dw_bwd_tc": 12,
              "dgnn_gather_tc_bwd": 13, "dgnn_layer_fwd": 14, "dgnn_layer_fwd_tc": 14, "dgnn_dense_bwd": 9, "dgnn_dw_bwd": 12,
              "dgnn_gather_bwd": 13}
_TRAFFIC_KERNELS = {"dgnn_gather_tc_fwd": [r"gather_tc_kernel<\d+, 0\b"],
                    "dgnn_gather_tc_bwd": [r"gather_tc_kernel<\d+, 1\b", r"dwe_tc_kernel<"],
                    "dgnn_dense_fwd_tc": [r"layer_tc_kernel<0, 0"], "dgnn_dense_bwd_tc": [r"layer_tc_kernel<2, 0"],
                    "dgnn_dw_bwd_tc": [r"dw_tc_kernel"]}


def launch_cells(name, args):
    i = _CELLS_ARG.get(name)
    return args[i] if i is not None else 0


def traffic_per_cell(name, tag):
    """(DRAM bytes per cell, source file) of the kernels behind one C-ABI call from the newest committed ncu capture."""
    import csv, glob, re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_layer_kernels.csv")))
    pats = _TRAFFIC_KERNELS.get(name)
    if not files or not pats:
        return None, None
    rows = list(csv.DictReader(open(files[-1])))
    total, hit = 0.0, 0
    for pat in pats:
        for r in rows:
            layer = r["layer"]
            same = layer == tag or ("->" not in tag and layer.startswith(tag + "->"))
            if same and re.search(pat, r["Kernel Name"]):
                total += float(r["dram_bytes_per_cell"]); hit += 1
                break
    if hit != len(pats):
        return None, os.path.basename(files[-1])
    return total, os.path.basename(files[-1])


class KernelProfile:
    """Per-launch device time of every C-ABI call, measured with CUDA events on the launching
    stream (used after the timed region to find the dominant kernel and its achieved GB/s)."""

    def __init__(self):
        self.rows = []
        self.rows_cells = {}

    def install(self):
        from dgnn_b200 import _lib
        self._orig = _lib.call
        prof = self

        def timed_call(name, *args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prof._orig(name, *args)
            e1.record()
            prof.rows.append((name, args, e0, e1))

        for modname in ("dgnn_b200._lib", "dgnn_b200.engine", "dgnn_b200.graph", "dgnn_b200.runModel"):
            mod = sys.modules.get(modname)
            if mod is not None and hasattr(mod, "call"):
                setattr(mod, "call", timed_call)

    def uninstall(self):
        for modname in ("dgnn_b200._lib", "dgnn_b200.engine", "dgnn_b200.graph", "dgnn_b200.runModel"):
            mod = sys.modules.get(modname)
            if mod is not None and hasattr(mod, "call"):
                setattr(mod, "call", self._orig)

    def summary(self, n_steps, peak_gbs, peak_src):
        torch.cuda.synchronize()
        agg = {}
        for name, args, e0, e1 in self.rows:
            ms = e0.elapsed_time(e1)
            nbytes, tag = algorithmic_bytes(name, args)
            k = (name, tag)
            r = agg.setdefault(k, dict(ms=0.0, launches=0, bytes=0, known=nbytes is not None))
            r["ms"] += ms; r["launches"] += 1; r["bytes"] += nbytes or 0
            self.rows_cells[k] = launch_cells(name, args)
        total = sum(r["ms"] for r in agg.values())
        top = max((k for k in agg if agg[k]["known"]), key=lambda k: agg[k]["ms"])
        r = agg[top]
        ach = r["bytes"] / (r["ms"] * 1e-3) / 1e9
        table = sorted(((k[0] + ":" + k[1], round(v["ms"] / n_steps, 4), v["launches"] // n_steps) for k, v in agg.items()),
                       key=lambda t: -t[1])
        # measured DRAM traffic of the dominant call: dram__bytes_read.sum + dram__bytes_write.sum per cell of its kernels in
        # the committed `ncu --set full` capture (profiles/r*_ncu_layer_kernels.csv, tools/ncu_kernels_csv.py), scaled to
        # the cells of this launch
        cells = int(self.rows_cells.get(top, 0))
        tpc, tsrc = traffic_per_cell(top[0], top[1])
        roof = {"bound": "hbm", "kernel": top[0] + ":" + top[1], "achieved": round(ach, 1), "peak": peak_gbs, "unit": "GB/s",
                "frac": round(ach / peak_gbs, 4), "peak_source": peak_src,
                "traffic": int(tpc * cells) if tpc and cells else None, "traffic_source": tsrc,
                "kernel_ms_per_launch": round(r["ms"] / r["launches"], 4),
                "kernel_share_of_step": round(r["ms"] / total, 4),
                "algorithmic_bytes_per_launch": r["bytes"] // r["launches"]}
        return roof, table[:int(os.environ.get("DGNN_BENCH_TABLE", "12"))], total / n_steps


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU reference arm


def cpu_reference(n_objects, steps, warmup, seed0=0):
    """fwd + loss + bwd + Adam of the oracle (plain PyTorch on the host cores) on ``n_objects``."""
    from oracle import trainer as otr
    from oracle.static_model import SurfaceNet as OracleNet
    from dgnn_b200.synthetic import make_clf, to_attr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d = make_objects(n_objects, seed0)
    data = batch_of(d, to_attr)
    clf = make_clf(convs=WIDTHS)
    torch.manual_seed(0)
    net = OracleNet(clf)
    opt = torch.optim.Adam(net.parameters(), lr=0.005)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        otr.train_step(net, opt, data, clf)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.sum(times))
    return d["n"] * steps / dt, cores, d["n"], dt / steps


def parity_vs_oracle(n_objects, dev, seed0=0):
    """First training step of the CUDA path against the CPU oracle on the cpu_baseline sample (same graphs, same initial
    weights): max |dz| / max(|z_ref|, mean|z_ref|) over the logits, the loss values, labels off ties."""
    from oracle import trainer as otr
    from oracle.static_model import SurfaceNet as OracleNet
    from dgnn_b200 import runModel as rm
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from dgnn_b200.synthetic import make_clf, to_attr
    d = make_objects(n_objects, seed0)
    data = batch_of(d, to_attr)
    torch.manual_seed(0)
    ref = OracleNet(make_clf(convs=WIDTHS)).train()
    clf = make_clf(convs=WIDTHS, device=str(dev))
    net = SurfaceNet(clf)
    net.load_state_dict(ref.state_dict(), strict=True)
    net.to(dev).train()
    with torch.no_grad():
        zr = ref(data)
        lr, _, _ = otr.cell_loss(zr, data.all.y, data.all.x[:, 0])
        z = net(data)
        loss = rm.cell_loss(z, data.all.y, data.all.x, clf)
    zr = zr.double(); z = z.cpu().double()
    scale = torch.maximum(zr.abs(), zr.abs().mean())
    ties = (zr[:, 0] - zr[:, 1]).abs() <= 2e-4 * zr.abs().mean()
    flips = int((((z[:, 1] > z[:, 0]) != (zr[:, 1] > zr[:, 0])) & ~ties).sum())
    return {"logits_max_rel": float(((z - zr).abs() / scale).max()), "loss_cuda": float(loss.item()), "loss_oracle": float(lr.item()),
            "label_flips_off_ties": flips, "cells": int(d["n"]), "tolerance": 1e-4}


def side_configs(net, dev, peak, args):
    """Single-GPU measurements of the other BASELINE configs on their own graphs (N = 1 only):
    cfg1  reconbench.yaml inference on ONE ~50 k-point scan-like graph (~330 k cells) + the CPU oracle on the same graph,
    cfg3  synthetic_room-shaped scene: scipy Delaunay of --cfg3-points random points (1 M -> ~6.7 M cells), inference,
    wide  the configs[1] training step at configs/modelnet.yaml:56 widths [128,256,512,1024] (tensor-pipe bound)."""
    from dgnn_b200 import runModel as rm, synthetic as og
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from dgnn_b200.synthetic import make_clf, to_attr
    out = {}

    def graph(points):
        adj, infinite, cen, _ = og.delaunay_graph(points)
        n = infinite.shape[0]
        x, ea, y = og.synthetic_features(n, infinite, seed=5)
        return to_attr(dict(x=torch.from_numpy(x), edge_attr=torch.from_numpy(ea), y=torch.from_numpy(y),
                            edge_index=torch.from_numpy(adj.T.astype(np.int64)).contiguous(),
                            pos=torch.from_numpy(cen.astype(np.float32)))), n

    def infer_ms(d, iters):
        dd = to_attr({k: v.to(dev) for k, v in d.items()})
        net.eval()
        with torch.no_grad():
            ms = _timed_region(lambda: net.inference_layer(dd), iters, 3, dev, 1) / iters
        net.train()
        return ms

    # cfg1
    d1, n1 = graph(og.scan_like_points(50_000, seed=0))
    ms1 = infer_ms(d1, 20)
    out["cfg1"] = {"cells": n1, "ms": round(ms1, 4), "cells_per_s": n1 / (ms1 * 1e-3),
                   "hbm_frac": round(n1 * 4024 / (ms1 * 1e-3) / 1e9 / peak, 4),
                   "workload": "configs[0] reconbench.yaml StaticEdgeFilters inference_layer, one 50 k-point scan-like graph, device resident"}
    if not args.no_cpu_baseline:
        from oracle.static_model import SurfaceNet as OracleNet
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ref = OracleNet(make_clf(convs=WIDTHS))
        ref.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
        ref.eval()
        with torch.no_grad():
            zr = ref.inference_layer(d1)
            t0 = time.perf_counter()
            for _ in range(2):
                ref.inference_layer(d1)
            cpu_s = (time.perf_counter() - t0) / 2
            net.eval()
            z = net.inference_layer(to_attr({k: v.to(dev) for k, v in d1.items()})).cpu()
            net.train()
        scale = torch.maximum(zr.abs(), zr.abs().mean())
        out["cfg1"]["cpu_baseline"] = {"value": n1 / cpu_s, "unit": "cells/s", "cores": cores, "kind": "port",
                                       "sample": "the same graph, 2 passes after 1 warm-up, oracle (plain PyTorch CPU)"}
        out["cfg1"]["parity_max_rel"] = float(((z - zr).abs() / scale).max())
    del d1
    # cfg3
    t0 = time.perf_counter()
    d3, n3 = graph(og.random_points(args.cfg3_points, seed=0))
    gen_s = time.perf_counter() - t0
    ms3 = infer_ms(d3, 5)
    out["cfg3"] = {"cells": n3, "points": args.cfg3_points, "ms": round(ms3, 3), "cells_per_s": n3 / (ms3 * 1e-3),
                   "hbm_frac": round(n3 * 4024 / (ms3 * 1e-3) / 1e9 / peak, 4), "host_delaunay_s": round(gen_s, 1),
                   "workload": "configs[2] synthetic_room-shaped scene, scipy Delaunay of random points, inference_layer, device resident"}
    del d3
    # wide widths: the configs[1] batch at modelnet.yaml widths
    import gc
    gc.collect(); torch.cuda.empty_cache()
    wide = (128, 256, 512, 1024)
    host = make_objects(16, seed0=0)
    clfw = make_clf(convs=wide, device=str(dev))
    torch.manual_seed(0)
    netw = SurfaceNet(clfw).to(dev).train()
    optw = rm.Adam(netw.parameters(), lr=0.005)
    dres = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    data = batch_of(dres, to_attr)

    def stepw():
        loss = rm.cell_loss(netw(data), data.all.y, data.all.x, clfw)
        optw.zero_grad(set_to_none=True)
        loss.backward()
        optw.step()

    msw = _timed_region(stepw, 5, 3, dev, 1) / 5
    flops = 0
    ws = [F0] + list(wide)
    for a, b in zip(ws[:-1], ws[1:]):
        flops += 4 * a * b + 8 * FE * a + 8 * a
    flops += 2 * wide[-1] * (wide[-1] // 2) + 4 * (wide[-1] // 2)
    out["wide"] = {"widths": [F0] + list(wide), "cells": host["n"], "ms_per_step": round(msw, 3),
                   "cells_per_s": host["n"] / (msw * 1e-3),
                   "algorithmic_tflops": round(3 * flops * host["n"] / (msw * 1e-3) / 1e12, 2),
                   "note": "configs/modelnet.yaml:56 widths, 16 objects, fwd+loss+bwd+Adam; tensor-pipe bound: 3 x forward FLOPs per "
                           "cell counted once (the 3xTF32 split issues 3 MMAs per product on top)"}
    return out


# --------------------------------------------------------------------------- partitioned scene (N > 1)

SCENE_DIMS_WEAK = {1: (128, 128, 256), 2: (128, 256, 256), 4: (256, 256, 256), 8: (256, 256, 512)}   # 8.39 M cells per GPU
SCENE_DIMS_STRONG = (256, 256, 512)                                                                 # 67.1 M cells
SCENE_DIMS_PARITY = (32, 32, 32)                                                                    # 65 536 cells


def _stage(msg):
    if os.environ.get("DGNN_BENCH_DEBUG"):
        print("[bench rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


def _maxr(v, dev, world):
    import torch.distributed as dist
    t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _timed_region(fn, steps, warmup, dev, world):
    """W warm-up + K timed calls between barrier + synchronize, CUDA events, max over ranks (ms for the K calls)."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    return _maxr(e0.elapsed_time(e1), dev, world)


def scene_inference(net, dims, rank, world, dev, iters=5, overlap=True):
    """Eval-mode partitioned inference of the whole lattice scene `dims` (sharded build); cells/s of the whole job."""
    from dgnn_b200 import scene as sc
    from dgnn_b200.partition import PartitionedInference
    n = 2 * dims[0] * dims[1] * dims[2]
    shard = sc.lattice_scene(dims, rank, world, dev)
    pi = PartitionedInference(net)
    g, maps, x0, ids, comm = pi.prepare_scene(shard, overlap=overlap)
    del shard
    net.eval()
    ms = _timed_region(lambda: pi.run(), iters, 2, dev, world) / iters
    ex_ms = 0.0
    if world > 1:            # the halo exchange of one 128-wide layer alone (pack + all-to-all), not overlapped
        h = torch.zeros((g.n_src, 128), dtype=torch.float32, device=dev)
        ex_ms = _timed_region(lambda: comm.exchange(h), 10, 2, dev, world) / 10
        del h
    res = {"cells": n, "dims": list(dims), "ms": round(ms, 3), "cells_per_s": n / (ms * 1e-3),
           "cells_per_gpu": maps.n_own, "halo_fraction": round(_maxr(maps.n_halo / max(maps.n_own, 1), dev, world), 5),
           "boundary_fraction": round(_maxr(maps.n_boundary / max(maps.n_own, 1), dev, world), 5),
           "exchange_ms_per_128wide_layer": round(ex_ms, 4), "exchanges_per_pass": 3,
           "limiting_collective": "all_to_all_single of boundary rows (NCCL), %.3f ms of %.3f ms per pass when not hidden"
                                  % (3 * ex_ms, ms),
           "overlap": bool(overlap and world > 1)}
    net.train()
    return res


UPD_DIMS = {1: (64, 128, 128), 2: (128, 128, 128), 4: (128, 128, 256), 8: (128, 256, 256)}   # 2.1 M cells per GPU


def updated_training(rank, world, dev, steps=5, warmup=2):
    """BASELINE configs[3] (cfg4): the Updated-edge-filter model (model_params [64,128,128,128], "sage+" head) trained on
    ONE lattice scene partitioned over the ranks (2.1 M cells per GPU: the variant materialises its edge state, 4 x F_in
    floats per cell and layer), fwd + kl loss + bwd + Adam, halo exchange forward, reverse halo exchange of the source
    gradients backward, gradient all-reduce.  cells/s of the whole job."""
    from dgnn_b200 import runModel as rm, scene as sc
    from dgnn_b200.partition import PartitionedUpdatedTraining
    from dgnn_b200.surfaceNetUpdatedEdgeFilters import SurfaceNet as UpdNet
    from dgnn_b200.synthetic import make_clf, to_attr
    dims = UPD_DIMS.get(world) or UPD_DIMS[8]
    n = 2 * dims[0] * dims[1] * dims[2]
    clf = to_attr(dict(training=dict(model_params=list(WIDTHS), model_name="sage+"),
                       features=dict(normalization_feature=0, keep_normalization_feature=0), temp=dict(device=str(dev))))
    loss_clf = make_clf(device=str(dev))
    torch.manual_seed(0)
    net = UpdNet(F0, clf).to(dev).train()
    if world > 1:
        import torch.distributed as dist
        for t in net.parameters():
            dist.broadcast(t.data, 0)
    opt = rm.Adam(net.parameters(), lr=0.005)
    shard = sc.lattice_scene(dims, rank, world, dev)
    pt = PartitionedUpdatedTraining(net)
    pt.prepare_scene(shard)
    y, w = shard.y, shard.w

    def step():
        _, logits = pt.forward()
        loss = rm.cell_loss(logits, y, w, loss_clf, group=None, distributed=world > 1)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        pt.allreduce_gradients()
        opt.step()

    ms = _timed_region(step, steps, warmup, dev, world) / steps
    return {"cells": n, "dims": list(dims), "ms_per_step": round(ms, 3), "cells_per_s": n / (ms * 1e-3), "model": "Updated edge "
            "filters, model_params %s, sage+ head, edge state materialised" % (list(WIDTHS),)}


def partition_parity(net, rank, world, dev):
    """max |partitioned - single GPU| over the logits of a small lattice scene: the partitioned path builds from shards
    (dgnn_b200.scene.lattice_scene), the single-GPU path from the whole graph through the ordinary loader layout."""
    import torch.distributed as dist
    from dgnn_b200 import scene as sc
    from dgnn_b200.partition import PartitionedInference
    from dgnn_b200.synthetic import to_attr
    dims = SCENE_DIMS_PARITY
    net.eval()
    pi = PartitionedInference(net)
    pi.prepare_scene(sc.lattice_scene(dims, rank, world, dev))
    ids, out = pi.run()
    n = 2 * dims[0] * dims[1] * dims[2]
    full = torch.zeros((n, out.shape[1]), dtype=torch.float32, device=dev)
    full[ids] = out
    if world > 1:
        dist.all_reduce(full)                       # every cell is owned by exactly one rank
    err = None
    if rank == 0:
        g = sc.lattice_global(dims)
        with torch.no_grad():
            ref = net.inference_layer(to_attr(g))
        err = float((full - ref).abs().max().item())
    net.train()
    return err


def run_partitioned(args, rank, world, local_rank):
    """N > 1: the partitioned-scene benchmark (module docstring)."""
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from dgnn_b200 import _lib, runModel as rm, scene as sc
    from dgnn_b200.partition import PartitionedTraining
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from dgnn_b200.synthetic import make_clf, to_attr

    launches = {"n": 0}
    orig_call = _lib.call
    mods = ("dgnn_b200._lib", "dgnn_b200.engine", "dgnn_b200.graph", "dgnn_b200.runModel", "dgnn_b200.partition")

    def counting_call(name, *a):
        launches["n"] += 1
        return orig_call(name, *a)

    def set_call(fn):
        for modname in mods:
            mod = sys.modules.get(modname)
            if mod is not None and hasattr(mod, "call"):
                setattr(mod, "call", fn)

    clf = make_clf(convs=WIDTHS, device=str(dev))
    torch.manual_seed(0)
    net = SurfaceNet(clf).to(dev).train()
    for t in list(net.parameters()) + list(net.buffers()):
        if t.is_floating_point():
            dist.broadcast(t.data, 0)
    opt = rm.Adam(net.parameters(), lr=0.005)

    # parity first (small scene): partitioned logits against the single-GPU path
    _stage("parity")
    parity = partition_parity(net, rank, world, dev)
    _stage("parity done %r" % (parity,))

    # ---- value: training step on the weak-scaled scene ------------------------------------------------------
    dims = SCENE_DIMS_WEAK.get(world) or SCENE_DIMS_WEAK[8]
    n_total = 2 * dims[0] * dims[1] * dims[2]
    shard = sc.lattice_scene(dims, rank, world, dev, need_backward=True)
    pt = PartitionedTraining(net)
    g, maps, x0, ids, comm = pt.prepare_scene(shard)
    host_feats = (shard.x.cpu().pin_memory(), shard.y.cpu().pin_memory(), shard.w.cpu().pin_memory())
    order = maps.order
    del shard
    set_call(counting_call)

    def step():
        _, logits = pt.forward()
        loss = pt.loss(logits)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        pt.allreduce_gradients()
        opt.step()
        return loss

    _stage("scene ready: own %d halo %d boundary %d" % (maps.n_own, maps.n_halo, maps.n_boundary))
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            step()
        launches["n"] = 0
        comm.stats = {"exchanges": 0, "bytes_sent": 0}
        ms = _timed_region(step, args.steps, 0, dev, world)
    n_launch = launches["n"]
    ex_per_step = comm.stats["exchanges"] / max(args.steps, 1)
    ex_bytes = comm.stats["bytes_sent"] / max(args.steps, 1)
    value = n_total * args.steps / (ms * 1e-3)

    # ---- e2e: node features + supervision of the step come from pinned HOST buffers (uploaded one step ahead on a
    # copy stream), the loss is read back; topology and edge attributes stay resident (one scene, many steps)
    copy_stream = torch.cuda.Stream(device=dev)
    n_own = maps.n_own
    h2d = sum(t.numel() * t.element_size() for t in host_feats)
    staged = {}

    def upload():
        with torch.cuda.stream(copy_stream):
            staged["t"] = tuple(t.to(dev, non_blocking=True) for t in host_feats)
            staged["ev"] = torch.cuda.Event(); staged["ev"].record(copy_stream)

    upload()

    def e2e_step():
        torch.cuda.current_stream().wait_event(staged["ev"])
        xs, ys, ws = staged["t"]
        for t in staged["t"]:
            t.record_stream(torch.cuda.current_stream())
        upload()                                            # next step's inputs while this step computes
        o = order if order is not None else slice(None)
        x0[:n_own, :xs.shape[1]] = xs[o]
        comm.exchange(x0)                                   # halo rows of the fresh features
        pt._sup = (ys[o], ws[o])
        return step().item()

    _stage("value %.3e" % value)
    e2e_steps = max(3, args.steps // 4)
    ms_e2e = _timed_region(e2e_step, e2e_steps, 1, dev, world)
    e2e_value = n_total * e2e_steps / (ms_e2e * 1e-3)
    _stage("e2e %.3e" % e2e_value)

    # ---- per-kernel times of one step on this rank -> roofline of the dominant kernel
    peak, peak_src = peak_hbm()
    set_call(orig_call)
    prof = KernelProfile()
    prof.install()
    for _ in range(2):
        step()
    prof.uninstall()
    roofline, table, kernel_ms = prof.summary(2, peak, peak_src)

    # ---- partitioned training against the single-GPU step on the small scene (loss + gradients)
    _stage("profile done")
    train_par = partition_train_parity(clf, rank, world, dev)
    _stage("train parity %r" % (train_par,))

    # free the training scene before the big inference scene
    del pt, g, x0, comm, staged
    opt.zero_grad(set_to_none=True)
    import gc
    gc.collect(); torch.cuda.empty_cache()
    infer = scene_inference(net, SCENE_DIMS_STRONG, rank, world, dev)
    infer["hbm_frac_per_gpu"] = round(infer["cells_per_s"] / world * 4024 / 1e9 / peak, 4)
    gc.collect(); torch.cuda.empty_cache()
    infer_noov = scene_inference(net, SCENE_DIMS_STRONG, rank, world, dev, iters=3, overlap=False)
    infer["ms_without_overlap"] = infer_noov["ms"]
    gc.collect(); torch.cuda.empty_cache()
    _stage("scene inference %r" % (infer,))

    upd = updated_training(rank, world, dev)
    gc.collect(); torch.cuda.empty_cache()
    _stage("updated training %r" % (upd,))

    # ---- the round-1 data-parallel object-batch number, as an extra key
    dp = dp_objects(args, rank, world, dev, clf)

    if rank == 0:
        config = {"workload": "configs[3]/[4]-shaped partitioned scene: ONE %d-cell graph (analytic diamond-cubic lattice %s, "
                              "4-regular, random features of the feat tool's shape) split into %d Morton ranges, per-layer halo "
                              "exchange over NCCL, Static kf96 model, fwd+loss+bwd+Adam, BN train mode with all-reduced "
                              "statistics, kl loss" % (n_total, "x".join(map(str, dims)), world),
                  "widths": [F0] + list(WIDTHS), "decoder": "128->64->2", "edge_features": FE,
                  "cells_per_gpu": maps.n_own, "parallelism": "graph partition x%d (halo all-to-all + gradient all-reduce)" % world,
                  "l2_policy": "inputs + saved activations per step (>20 GB per GPU) exceed the 126 MB L2"}
        out = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "cells_per_gpu_per_step": maps.n_own,
               "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                       "ms_per_step": ms_e2e / e2e_steps,
                       "resident": "topology and edge attributes of the scene (loaded once); node features, targets and "
                                   "loss weights are uploaded every step"},
               "gpu_launches": int(n_launch), "clocks": clocks.summary(), "roofline": roofline,
               "step_hbm_frac": round(maps.n_own * 19_900 / (ms / args.steps * 1e-3) / 1e9 / peak, 4),
               "kernel_ms_per_step": round(kernel_ms, 3),
               "scene": {"cells": n_total, "dims": list(dims), "generator": "diamond-cubic lattice, sharded on-device build",
                         "halo_fraction": round(maps.n_halo / max(maps.n_own, 1), 5),
                         "boundary_fraction": round(maps.n_boundary / max(maps.n_own, 1), 5),
                         "halo_exchanges_per_step": ex_per_step, "halo_bytes_sent_per_gpu_per_step": int(ex_bytes)},
               "scene_inference": infer,
               "partition_vs_single_max_abs": parity,
               "partition_train_vs_single": train_par,
               "updated_training": upd,
               "dp": dp, "kernels": table, "cpu_baseline": None}
        print(json.dumps(out), flush=True)
    _stage("done")
    dist.barrier()
    dist.destroy_process_group()


def partition_train_parity(clf, rank, world, dev):
    """One training step (loss + every gradient) of the partitioned small scene against the single-GPU step."""
    import torch.distributed as dist
    from dgnn_b200 import runModel as rm, scene as sc
    from dgnn_b200.partition import PartitionedTraining
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from dgnn_b200.synthetic import to_attr
    dims = SCENE_DIMS_PARITY
    torch.manual_seed(1)
    net = SurfaceNet(clf).to(dev).train()
    for t in list(net.parameters()) + list(net.buffers()):
        if t.is_floating_point() and world > 1:
            dist.broadcast(t.data, 0)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    pt = PartitionedTraining(net)
    pt.prepare_scene(sc.lattice_scene(dims, rank, world, dev, need_backward=True))
    _, logits = pt.forward()
    loss = pt.loss(logits)
    loss.backward()
    pt.allreduce_gradients()
    res = None
    if rank == 0:
        ref = SurfaceNet(clf).to(dev).train()
        ref.load_state_dict(state)
        g = to_attr(sc.lattice_global(dims))
        n = g.x.shape[0]
        ei = g.edge_index
        batch = to_attr(dict(all=g, batch_n_id=torch.arange(n), batch_adjs=[(ei, torch.arange(ei.shape[1]), (n, n))] * 5))
        z = ref(batch)
        lr = rm.cell_loss(z, g.y, g.x, clf)
        lr.backward()
        worst = 0.0
        refp = dict(ref.named_parameters())
        for k, p in net.named_parameters():
            gr = refp[k].grad
            if gr is None or float(gr.norm()) < 1e-6:
                continue
            worst = max(worst, float((p.grad - gr).norm() / gr.norm()))
        res = {"loss_abs_diff": abs(float(loss.item()) - float(lr.item())), "grad_max_rel_frobenius": worst}
    return res


def dp_objects(args, rank, world, dev, clf):
    """Round-1 multi-GPU mode kept as an extra key: data-parallel over object batches (configs[1] per rank), one
    gradient all-reduce per step."""
    import torch.distributed as dist
    from dgnn_b200 import runModel as rm
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from dgnn_b200.synthetic import to_attr
    host = make_objects(args.objects, seed0=1000 * rank)
    torch.manual_seed(0)
    net = SurfaceNet(clf).to(dev).train()
    opt = rm.Adam(net.parameters(), lr=0.005)
    params = list(net.parameters())
    dres = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    data = batch_of(dres, to_attr)

    def step():
        loss = rm.cell_loss(net(data), data.all.y, data.all.x, clf)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(flat)
        flat /= world
        off = 0
        for p in params:
            p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
        opt.step()

    steps = max(5, args.steps // 2)
    ms = _timed_region(step, steps, 3, dev, world)
    return {"value": world * host["n"] * steps / (ms * 1e-3), "unit": "cells/s", "ms_per_step": ms / steps,
            "cells_per_gpu_per_step": host["n"], "note": "data-parallel object batches (weak scaling), gradient all-reduce only"}


# --------------------------------------------------------------------------- main


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=64, help="object graphs per GPU per step")
    ap.add_argument("--cpu-objects", type=int, default=4, help="objects in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true", help="skip the cfg1 / cfg3 / wide-width side measurements")
    ap.add_argument("--cfg3-points", type=int, default=1_000_000, help="points of the cfg3 scene (scipy Delaunay on the host)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "configs[1] ModelNet10-shaped training: %d object graphs/GPU/step x ~19.7k cells "
                          "(3000 scan-like points each), fwd+loss+bwd+Adam, BN train mode, kl loss" % args.objects,
              "widths": [F0] + list(WIDTHS), "decoder": "128->64->2", "edge_features": FE,
              "objects_per_gpu": args.objects, "parallelism": "dp%d (gradient all-reduce)" % world if world > 1 else "single GPU",
              "l2_policy": "inputs + saved activations per step (>1 GB) exceed the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        warm = max(1, min(args.warmup, 1))
        v, cores, n_cells, s_per = cpu_reference(args.cpu_objects, steps, warm)
        sample = "%d objects (%d cells) per step, %d steps after %d warm-up" % (args.cpu_objects, n_cells, steps, warm)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": s_per * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if world > 1:
        run_partitioned(args, rank, world, local_rank)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from dgnn_b200 import _lib, runModel as rm
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from dgnn_b200.synthetic import make_clf, to_attr  # clf / attr-dict stand-ins for Munch

    # count kernel launches issued through the C ABI
    launches = {"n": 0}
    orig_call = _lib.call
    two = {"dgnn_ell_build": 2}

    def counting_call(name, *a):
        launches["n"] += two.get(name, 1)
        return orig_call(name, *a)

    for modname in ("dgnn_b200._lib", "dgnn_b200.engine", "dgnn_b200.graph", "dgnn_b200.runModel"):
        setattr(sys.modules[modname], "call", counting_call)

    host = make_objects(args.objects, seed0=1000 * rank)
    n_cells = host["n"]
    clf = make_clf(convs=WIDTHS, device=str(dev))
    torch.manual_seed(0)
    net = SurfaceNet(clf).to(dev).train()
    opt = rm.Adam(net.parameters(), lr=0.005)
    params = [p for p in net.parameters()]

    # device-resident batch (value) ---------------------------------------------------------
    dres = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    data = batch_of(dres, to_attr)

    def step(batch, d_all):
        logits = net(batch)
        loss = rm.cell_loss(logits, d_all.y, d_all.x, clf)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat /= world
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
        opt.step()
        return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches["n"] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches["n"]

    # `value`: the step captured once into a CUDA graph (runModel.GraphedStep) and replayed - same kernels, same
    # arithmetic (bit-identical, tests/test_gpu_scale.py), one graph launch instead of ~140 kernel launches per step.
    # The eager number is kept beside it.
    ms_eager, n_launch = timed(lambda: step(data, data.all), max(5, args.steps // 2), args.warmup)
    ms_eager /= max(5, args.steps // 2)
    gstep = rm.GraphedStep(lambda: rm.cell_loss(net(data), data.all.y, data.all.x, clf), opt, warmup=0)
    n_launch = n_launch // max(5, args.steps // 2) * args.steps      # kernels inside the timed region (replayed, not re-launched)
    with ClockSampler(local_rank) as clocks:
        ms, _ = timed(gstep, args.steps, args.warmup)
    value = world * n_cells * args.steps / (ms * 1e-3)

    # end to end through the public API with HOST buffers (e2e) ------------------------------
    # every step uploads its own inputs (x, edge_attr, edge_index, y, pos + the batch's n_id / e_id) from pinned host
    # memory; the upload of step i+1 runs on a copy stream while step i computes (runModel.BatchPrefetcher), the
    # graph layout is rebuilt from the uploaded tensors every step (no cached plan), the loss is read back.
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    net.cache_graphs = False
    hb = batch_of(pinned, to_attr)              # host batch, as the reference trainer holds it
    hb.batch_n_id = hb.batch_n_id.pin_memory()
    e_id_pinned = hb.batch_adjs[0][1].pin_memory()
    hb.batch_adjs = [(a[0], e_id_pinned, a[2]) for a in hb.batch_adjs]
    h2d = sum(t.numel() * t.element_size() for t in (pinned["x"], pinned["edge_attr"], pinned["y"], pinned["edge_index"],
                                                     pinned["pos"], hb.batch_n_id, e_id_pinned))
    pf = rm.BatchPrefetcher(dev)
    pf.put(hb)

    e2e_steps = max(3, args.steps // 4)
    e2e_state = {"calls": 0, "pending": None, "losses": []}

    def e2e_step():
        cur = pf.get()
        pf.put(hb)                              # the next step's upload overlaps this step's kernels
        loss = step(cur, cur.all)
        # device -> host read of every step's result, one step late: the loss of step i-1 is read while step i is queued
        # (a training loop that logs its loss does the same; a blocking read right here would idle the GPU for the ~150
        # launches of the next step).  The last timed call reads its own loss too, inside the timed region.
        prev, e2e_state["pending"] = e2e_state["pending"], loss
        e2e_state["calls"] += 1
        if prev is not None:
            e2e_state["losses"].append(prev.item())
        if e2e_state["calls"] == 2 + e2e_steps:
            e2e_state["losses"].append(loss.item())

    ms_e2e, _ = timed(e2e_step, e2e_steps, 2)
    assert len(e2e_state["losses"]) == 2 + e2e_steps and all(math.isfinite(v) for v in e2e_state["losses"])
    e2e_value = world * n_cells * e2e_steps / (ms_e2e * 1e-3)
    net.cache_graphs = True

    # per-kernel device time of one step -> roofline of the dominant kernel --------------------
    peak, peak_src = peak_hbm()
    prof = KernelProfile()
    for modname in ("dgnn_b200._lib", "dgnn_b200.engine", "dgnn_b200.graph", "dgnn_b200.runModel"):
        setattr(sys.modules[modname], "call", orig_call)
    prof.install()
    n_prof = 3
    for _ in range(n_prof):
        step(data, data.all)
    prof.uninstall()
    roofline, table, kernel_ms = prof.summary(n_prof, peak, peak_src)

    # secondary: eval-mode whole-batch inference (configs[0]/[2] shape of work), device resident
    net.eval()
    dall = to_attr({k: v for k, v in dres.items() if k != "n"})
    with torch.no_grad():
        for _ in range(3):
            net.inference_layer(dall)
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(10):
            net.inference_layer(dall)
        i1.record()
        torch.cuda.synchronize()
    infer_ms = i0.elapsed_time(i1) / 10
    net.train()

    # the fixed 67.1 M-cell scene of the multi-GPU runs on ONE GPU (strong-scaling reference point of `scene_inference`)
    del data, dres, dall, hb, pinned, pf
    opt.zero_grad(set_to_none=True)
    import gc
    gc.collect(); torch.cuda.empty_cache()
    try:
        scene_inf = scene_inference(net, SCENE_DIMS_STRONG, 0, 1, dev, iters=3)
        scene_inf["hbm_frac_per_gpu"] = round(scene_inf["cells_per_s"] * 4024 / 1e9 / peak, 4)
    except torch.OutOfMemoryError:
        scene_inf = {"error": "out of memory on one GPU"}
    gc.collect(); torch.cuda.empty_cache()
    part_parity = partition_parity(net, 0, 1, dev)       # sharded build (1 shard) against the ordinary loader path
    upd = updated_training(0, 1, dev)
    gc.collect(); torch.cuda.empty_cache()
    side = {}
    if not args.no_side_configs:
        side = side_configs(net, dev, peak, args)
        gc.collect(); torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, nc, s_per = cpu_reference(args.cpu_objects, 2, 1)
        cpu = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
               "sample": "%d objects (%d cells) per step, 2 steps after 1 warm-up, oracle (plain PyTorch CPU)" % (args.cpu_objects, nc)}
        parity = parity_vs_oracle(args.cpu_objects, dev)

    # whole-step algorithmic bytes (kf96 training fwd+bwd ~ 19.9 KB/cell)
    step_bytes_per_cell = 19_900
    out = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "cells_per_gpu_per_step": n_cells,
           "launch": "CUDA graph replay of the captured step (runModel.GraphedStep)", "ms_per_step_eager": ms_eager,
           "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                   "ms_per_step": ms_e2e / e2e_steps},
           "gpu_launches": int(n_launch),
           "clocks": clocks.summary(),
           "roofline": roofline,
           "step_hbm_frac": round(n_cells * step_bytes_per_cell / (ms / args.steps * 1e-3) / 1e9 / peak, 4),
           "kernel_ms_per_step": round(kernel_ms, 3),
           "inference": {"cells_per_s_per_gpu": n_cells / (infer_ms * 1e-3), "ms": round(infer_ms, 3),
                         "hbm_frac": round(n_cells * 4024 / (infer_ms * 1e-3) / 1e9 / peak, 4),
                         "note": "eval-mode inference_layer on the same batch, device resident, 4 024 B/cell"},
           "scene_inference": scene_inf,
           "updated_training": upd,
           "cfg1_inference": side.get("cfg1"), "cfg3_inference": side.get("cfg3"), "wide_training": side.get("wide"),
           "partition_vs_single_max_abs": part_parity,
           "parity": parity,
           "parity_max_rel": parity["logits_max_rel"] if parity else None,
           "kernels": table,
           "cpu_baseline": cpu}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
