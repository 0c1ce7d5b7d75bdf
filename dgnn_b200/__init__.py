"""dgnn_b200 — B200-native (sm_100a) implementation of the DGNN cell-classification hot path.

Modules mirror the reference files they replace:
  surfaceNetStaticEdgeFilters   learning/surfaceNetStaticEdgeFilters.py  (SurfaceNet, drop-in)
  surfaceNetUpdatedEdgeFilters  learning/surfaceNetUpdatedEdgeFilters.py (SAGEConv math + forward)
  runModel                      learning/runModel.py loss / regulariser / Adam step
  graph                         processing/data.py adjacency -> ELL-4 re-layout
The compute lives in csrc/libdgnn_b200.so (C ABI: include/dgnn_b200.h); there is no CPU fallback.
"""
from ._lib import DgnnError, LIB_PATH, lib  # noqa: F401

__all__ = ["DgnnError", "LIB_PATH", "lib"]
