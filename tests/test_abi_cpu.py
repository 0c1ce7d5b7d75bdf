"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes table matches the header, and the product refuses to run without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dgnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int|const char\*)\s+(dgnn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from dgnn_b200 import _lib
    l = _lib.lib()
    decl = header_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(l, name), name
    assert l.dgnn_version() == 100


def test_ctypes_table_matches_header():
    from dgnn_b200 import _lib
    decl = header_functions()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, n in decl.items():
        assert len(_lib.SIGNATURES[name]) == n, (name, n, len(_lib.SIGNATURES[name]))


def test_no_cpu_fallback():
    from dgnn_b200._lib import DgnnError
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from oracle.static_model import make_clf
    from tests.helpers import data_all, make_graph
    net = SurfaceNet(make_clf(device="cpu"))
    d = data_all(make_graph(60, 0))
    with pytest.raises(DgnnError):
        net.inference_layer(d)


def test_state_dict_layout_is_the_reference_layout(kf96_state):
    from dgnn_b200.surfaceNetStaticEdgeFilters import SurfaceNet
    from oracle.static_model import make_clf
    net = SurfaceNet(make_clf())
    net.load_state_dict(kf96_state, strict=True)
    assert list(net.state_dict().keys()) == list(kf96_state.keys())
    assert net.num_layers == 4 and net.output_dim == 2 and net.n_classes == 2


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dgnn_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
