"""ctypes binding of ``libdgnn_b200.so`` (the C ABI declared in ``include/dgnn_b200.h``).

There is no CPU or eager-PyTorch fallback: if the shared library is missing, or the device is
not an sm_100 GPU, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# DGNN_B200_LIB: development override (kernel variants built by tools/build_variant.sh)
LIB_PATH = os.environ.get("DGNN_B200_LIB") or os.path.join(_HERE, "csrc", "libdgnn_b200.so")

P = c_void_p
I = c_int
L = c_int64
F = c_float

# name -> argument types (all return int unless listed in _RET)
SIGNATURES = {
    "dgnn_version": [],
    "dgnn_last_error": [],
    "dgnn_device_check": [I],
    "dgnn_sm_count": [],
    "dgnn_reserve_sms": [I],
    "dgnn_small_grid": [],
    "dgnn_ell_from_adjacency": [P, L, P, P, P, P],
    "dgnn_ell_build": [P, P, L, L, L, I, P, P, P, P, P],
    "dgnn_morton_codes": [P, L, P, P, I, P, P],
    "dgnn_perm_apply_ell": [P, P, P, L, P, P],
    "dgnn_gather_rows": [P, P, L, I, P, P],
    "dgnn_scatter_rows": [P, P, L, I, P, P],
    "dgnn_add_rows": [P, P, L, I, P, P],
    "dgnn_sampler_degree": [P, L, P, P, P],
    "dgnn_sampler_expand": [P, L, P, P, P, P, P, P, P],
    "dgnn_sampler_set_loc": [P, L, I, P, P],
    "dgnn_sampler_mark": [P, L, P, P, P, P],
    "dgnn_sampler_assign": [P, L, P, P, L, P, P, P, P, P],
    "dgnn_column_moments": [P, L, I, I, I, P, P, I, P],
    "dgnn_column_affine": [P, L, I, I, I, P, P, I, P, P],
    "dgnn_column_moments_f64": [P, L, I, I, I, P, P, I, P],
    "dgnn_column_standardize_f64": [P, L, I, I, I, P, P, I, I, P, P],
    "dgnn_edge_relayout": [P, P, P, P, L, I, P, P, P],
    "dgnn_edge_relayout_idx": [P, P, P, P, P, L, I, P, P, P],
    "dgnn_layer_grid": [I, I],
    "dgnn_layer_fwd": [P, P, P, I, P, P, I, P, P, P, P, P, P, I, L, I, I, P, P, P, P],
    "dgnn_tc_supported": [I, I, I],
    "dgnn_tc_grid": [],
    "dgnn_tc_packed_floats": [I, I, I],
    "dgnn_tc_slice": [I],
    "dgnn_pack_b_tf32": [P, I, I, I, I, I, P, P],
    "dgnn_layer_fwd_tc": [P, P, P, I, P, P, I, P, P, P, P, P, P, I, L, I, I, P, P, P, P],
    "dgnn_dense_bwd_tc": [P, P, P, P, P, P, P, P, P, L, I, I, P, P, P, P],
    "dgnn_gather_tc_supported": [I, I],
    "dgnn_gather_tc_fwd": [P, P, P, I, P, P, I, P, P, L, I, P, P],
    "dgnn_gather_tc_bwd": [P, P, P, P, I, P, P, P, P, P, P, P, I, L, L, I, P, P, P, P],
    "dgnn_dense_fwd_tc": [P, P, P, P, I, P, P, P, P, I, L, I, I, P, P, P],
    "dgnn_dw_tc_supported": [I, I],
    "dgnn_dw_bwd_tc": [P, P, P, P, P, P, P, P, P, P, P, I, L, I, I, I, P, P, P],
    "dgnn_gather_phi_fwd": [P, P, P, I, P, P, L, I, P, P],
    "dgnn_upd_edge_bwd": [P, P, P, I, P, P, P, P, P, L, I, P, P],
    "dgnn_gather_phi_bwd": [P, P, P, P, P, P, I, L, L, I, P, P],
    "dgnn_relu_mask": [P, P, L, P, P],
    "dgnn_norm_finalize": [P, I, L, I, P, P, F, F, I, P, P, P, P, P, P, P],
    "dgnn_norm_eval_affine": [P, P, P, P, F, I, P, P, P],
    "dgnn_rowdot_fwd": [P, P, P, I, P, P, L, I, I, P, P],
    "dgnn_affine_relu": [P, P, P, I, L, I, P, P],
    "dgnn_kl_loss_fwd": [P, P, I, P, I, I, L, P, P],
    "dgnn_kl_loss_finalize": [P, I, P, P],
    "dgnn_kl_loss_bwd": [P, P, I, P, I, I, L, P, P, P, P],
    "dgnn_point_loss_fwd": [P, P, I, P, I, I, I, L, P, P],
    "dgnn_point_loss_bwd": [P, P, I, P, I, I, I, L, P, P, P, P],
    "dgnn_edge_reg_fwd": [P, P, P, L, P, P],
    "dgnn_edge_reg_bwd": [P, P, P, L, L, F, P, P, P, P],
    "dgnn_rowdot_bwd": [P, P, P, P, P, P, I, P, L, I, I, P, P, P],
    "dgnn_act_bwd": [P, P, P, P, P, P, I, L, I, P, P, P],
    "dgnn_reduce_partials": [P, I, I, P, P],
    "dgnn_norm_bwd_coeffs": [P, P, L, I, P, P, I, P, P, P, P],
    "dgnn_dense_bwd": [P, P, P, P, P, P, P, P, P, L, I, I, I, P, P, P, P],
    "dgnn_dw_splits": [I, I],
    "dgnn_dw_bwd": [P, P, P, P, P, P, P, P, P, P, P, I, L, I, I, I, P, P],
    "dgnn_reduce_partials_f32": [P, I, L, P, P],
    "dgnn_gather_bwd_grid": [I],
    "dgnn_gather_bwd": [P, P, P, P, I, P, P, P, P, P, P, P, I, L, L, I, P, P, P],
    "dgnn_edge_filter_bwd": [P, P, P, I, P, P, P, P, P, I, L, L, I, P, P],
    "dgnn_adam_step": [P, P, P, P, L, F, F, F, F, I, P],
    "dgnn_adam_multi": [P, I, L, F, F, F, F, I, P],
    "dgnn_adam_multi_dev": [P, I, L, P, P, P],
    "dgnn_gc_terminals": [P, L, F, P, P, P, P, P],
    "dgnn_gc_push_relabel": [L, P, P, P, P, P, P, I, I, P],
    "dgnn_gc_bfs_init": [L, P, P, I, P],
    "dgnn_gc_bfs_step": [L, P, P, P, P, I, I, P, P],
    "dgnn_gc_active": [L, P, P, I, P, P],
    "dgnn_gc_labels": [L, P, I, P, P],
    "dgnn_gc_energy_grid": [],
    "dgnn_gc_energy": [L, P, F, P, I, P, P, P],
    "dgnn_argmax_labels": [P, L, I, P, P],
    "dgnn_scores": [P, L, I, P, P, P],
    "dgnn_interface_facets": [P, L, P, L, P, P],
}
_RET = {"dgnn_last_error": c_char_p}

_lib = None


class DgnnError(RuntimeError):
    pass


def lib():
    """The loaded library; raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DgnnError(
                "dgnn_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the header and the library disagree
            fn.argtypes = args
            fn.restype = _RET.get(name, c_int)
        _lib = l
    return _lib


def call(name, *args):
    """Call an int-returning entry point and raise ``DgnnError`` with the library's message."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise DgnnError("%s failed: %s" % (name, lib().dgnn_last_error().decode()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_checked = set()


def check_device(index: int):
    if index not in _checked:
        call("dgnn_device_check", int(index))
        _checked.add(index)
