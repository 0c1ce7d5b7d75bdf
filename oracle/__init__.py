"""CPU oracle for the DGNN cell-classification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dgnn_b200/`` may import this package:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker.

It restates, in plain PyTorch / NumPy on the CPU, the algorithm of

* ``learning/surfaceNetStaticEdgeFilters.py``   (``oracle.static_model``)
* ``learning/surfaceNetUpdatedEdgeFilters.py``  (``oracle.updated_model``)
* ``learning/runModel.py`` loss / regulariser   (``oracle.trainer``)
* the facet-adjacency layout of ``processing/data.py`` (``oracle.graph``)

in the *edge-list / scatter* formulation of the reference, i.e. independent of
the ELL-4 formulation the CUDA path uses.

Parity pin: the reference has no tests or golden vectors,
and its arithmetic lives in un-vendored third-party code (torch_geometric 2.0.2,
torch_scatter 2.0.9).  The oracle is pinned instead against the reference's OWN
model / trainer / loader source executed unmodified in the build container over a
small shim of those third-party entry points (``tests/golden/make_golden.py``);
the resulting fixtures are committed under ``tests/golden/``.  See DESIGN.md.
"""
