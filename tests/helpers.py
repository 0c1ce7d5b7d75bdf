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


def grad_close(g_new, g_ref, rtol=1e-2):
    """Gradient agreement that is robust to single ReLU-mask flips (a pre-activation within ~1e-6 of
    zero may fall on the other side in fp32; one flip moves a gradient's norm by ~4e-3 while a wrong
    formula moves every element): relative Frobenius error <= rtol AND the median element-wise error
    <= 1e-3 of the gradient's RMS.  Returns (error, tolerance) with error > tolerance on failure."""
    g_new = g_new.detach().cpu().double(); g_ref = g_ref.detach().cpu().double()
    nr = float(g_ref.norm())
    if nr < 1e-6:                       # analytically-zero gradient (bias in front of a BatchNorm): both
        return float((g_new - g_ref).abs().max()), 2e-5   # sides are rounding noise of a cancelling sum
    frob = float((g_new - g_ref).norm()) / nr
    rms = nr / (g_ref.numel() ** 0.5)
    med = float((g_new - g_ref).abs().median()) / rms
    return max(frob / rtol, med / 1e-3), 1.0
