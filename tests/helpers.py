"""Shared test helpers (CPU side: oracle objects and tolerances)."""
import numpy as np
import torch

from oracle import graph as og
from oracle.static_model import to_attr

REL_TOL = 1e-4  # north_star: fp32 logits agree within 1e-4 relative (to the logit scale)


def logits_close(z_new, z_ref, tol=REL_TOL):
    """|z_new - z_ref| <= tol * max(|z_ref|, mean|z_ref|) element-wise."""
    z_new = np.asarray(z_new, dtype=np.float64)
    z_ref = np.asarray(z_ref, dtype=np.float64)
    scale = np.maximum(np.abs(z_ref), np.abs(z_ref).mean())
    err = np.abs(z_new - z_ref) / scale
    return float(err.max()), bool((err <= tol).all())


def labels_equal_off_ties(z_new, z_ref, tie=2e-4):
    """Labels (argmax, ties -> 0) identical except where |z0 - z1| <= tie * mean|z| in the oracle."""
    z_new = np.asarray(z_new); z_ref = np.asarray(z_ref)
    la, lb = og.labels_from_logits(z_new), og.labels_from_logits(z_ref)
    ties = np.abs(z_ref[:, 0] - z_ref[:, 1]) <= tie * np.abs(z_ref).mean()
    return int(((la != lb) & ~ties).sum()), int(ties.sum())


def make_graph(n_points, seed, scan_like=True):
    pts = og.scan_like_points(n_points, seed=seed) if scan_like else og.random_points(n_points, seed=seed)
    adj, infinite, cen, tets = og.delaunay_graph(pts)
    n = infinite.shape[0]
    x, ea, y = og.synthetic_features(n, infinite, seed=seed + 1)
    return dict(adj=adj, infinite=infinite, cen=cen, x=x, ea=ea, y=y, n=n)


def data_all(g, with_pos=False):
    d = dict(x=torch.from_numpy(g["x"]), edge_attr=torch.from_numpy(g["ea"]), y=torch.from_numpy(g["y"]),
             edge_index=torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous())
    if with_pos:
        d["pos"] = torch.from_numpy(g["cen"].astype(np.float32))
    return to_attr(d)


def full_batch(d, n_layers_plus=5):
    n = d.x.shape[0]
    ei = d.edge_index
    return to_attr(dict(all=d, batch_n_id=torch.arange(n),
                        batch_adjs=[(ei, torch.arange(ei.shape[1]), (n, n))] * n_layers_plus))


GRAD_FROB_TOL = 1e-2   # relative Frobenius error of a parameter gradient
GRAD_MED_TOL = 2e-3    # median element-wise error / RMS of the gradient


def grad_close(g_new, g_ref, rtol=GRAD_FROB_TOL):
    """THE gradient tolerance of this repo: relative Frobenius error <= 1e-2 AND median
    element-wise error <= 2e-3 of the gradient's RMS.  Why not 1e-4 like the logits: a pre-activation within the
    forward's rounding error of zero falls on the other side of the ReLU in another arithmetic, and each such mask flip
    moves whole rows of the downstream gradients.  The fp32 reference formulation ITSELF differs from its fp64 run by
    1e-4 .. 1e-3 (Frobenius) and up to 8e-4 (median, layer-0 parameters of the [128,256,512,1024] model); the CUDA path
    (3xTF32, forward error 7e-6 vs 2e-6) measures 1e-3 .. 3e-3 and up to 1.5e-3 (tools/diag_grad_wide.py on B200).  A
    wrong formula moves every element by O(1).  Returns (error, tolerance) with error > tolerance on failure."""
    g_new = g_new.detach().cpu().double(); g_ref = g_ref.detach().cpu().double()
    nr = float(g_ref.norm())
    if nr < 1e-6:                       # analytically-zero gradient (bias in front of a BatchNorm): both
        return float((g_new - g_ref).abs().max()), 2e-5   # sides are rounding noise of a cancelling sum
    frob = float((g_new - g_ref).norm()) / nr
    rms = nr / (g_ref.numel() ** 0.5)
    med = float((g_new - g_ref).abs().median()) / rms
    return max(frob / rtol, med / GRAD_MED_TOL), 1.0
