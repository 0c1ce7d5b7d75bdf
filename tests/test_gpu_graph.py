"""GPU: graph layout kernels bit-exact against the NumPy oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import graph as og
from tests.helpers import make_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return make_graph(1500, seed=5)


def test_ell_from_adjacency_bit_exact(g):
    from dgnn_b200.graph import build_full_graph
    ei = torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()
    ea = torch.from_numpy(g["ea"])
    eg = build_full_graph(ei, ea, g["n"], "cuda:0", order="none")
    nbr, rslot = og.ell_from_adjacency(g["adj"])
    assert np.array_equal(eg.nbr.cpu().numpy(), nbr)
    ea_in, ea_own = og.relayout_edges(g["ea"], nbr, rslot)
    assert np.array_equal(eg.ea_in.cpu().numpy(), ea_in)
    assert np.array_equal(eg.ea_own.cpu().numpy(), ea_own)
    assert eg.perm is None


def test_morton_permutation_and_relayout_bit_exact(g):
    from dgnn_b200.graph import build_full_graph
    ei = torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()
    pos = torch.from_numpy(g["cen"].astype(np.float32))
    eg = build_full_graph(ei, torch.from_numpy(g["ea"]), g["n"], "cuda:0", pos=pos, order="morton")
    perm = og.morton_perm(g["cen"])
    assert np.array_equal(eg.perm.cpu().numpy(), perm)
    assert np.array_equal(eg.inv.cpu().numpy(), og.invert_perm(perm))
    nbr, rslot = og.ell_from_adjacency(g["adj"])
    assert np.array_equal(eg.nbr.cpu().numpy(), og.apply_perm_ell(nbr, perm))
    ea_in, ea_own = og.relayout_edges(g["ea"], nbr, rslot, perm)
    assert np.array_equal(eg.ea_in.cpu().numpy(), ea_in)
    assert np.array_equal(eg.ea_own.cpu().numpy(), ea_own)
    x = torch.from_numpy(g["x"][:, 1:]).cuda()
    xp = eg.permute_rows(x)
    assert np.array_equal(xp.cpu().numpy(), g["x"][:, 1:][perm])
    assert torch.equal(eg.unpermute_rows(xp), x)


def test_rcm_order_is_a_permutation(g):
    from dgnn_b200.graph import build_full_graph
    ei = torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()
    eg = build_full_graph(ei, None, g["n"], "cuda:0", order="rcm")
    perm = eg.perm.cpu().numpy()
    assert np.array_equal(np.sort(perm), np.arange(g["n"]))
    nbr, _ = og.ell_from_adjacency(g["adj"])
    assert np.array_equal(eg.nbr.cpu().numpy(), og.apply_perm_ell(nbr, perm))


def test_generic_edge_list_builder_bit_exact(g):
    from dgnn_b200.graph import build_from_edges
    rng = np.random.default_rng(0)
    adj = g["adj"]
    keep = rng.random(adj.shape[0]) < 0.7          # ragged rows: 0..4 in-edges per target
    order = rng.permutation(int(keep.sum()))       # arbitrary edge order
    src = adj[keep, 0][order].astype(np.int64)
    tgt = adj[keep, 1][order].astype(np.int64)
    e_id = np.nonzero(keep)[0][order]
    n = g["n"]
    eg = build_from_edges(torch.from_numpy(np.stack([src, tgt])), torch.from_numpy(e_id), torch.from_numpy(g["ea"]),
                          n, n, "cuda:0")
    nbr, eid, cnt = og.ell_from_edges(src, tgt, n)
    assert np.array_equal(eg.nbr.cpu().numpy(), nbr)
    ea_rows = g["ea"][e_id]
    exp = np.where((eid >= 0)[:, :, None], ea_rows[np.maximum(eid, 0)], 0.0)
    assert np.array_equal(eg.ea_in.cpu().numpy(), exp)
    onbr, oeid, _ = og.ell_from_edges(tgt, src, n)
    assert np.array_equal(eg.onbr.cpu().numpy(), onbr)


def test_more_than_four_in_edges_is_rejected(g):
    from dgnn_b200._lib import DgnnError
    from dgnn_b200.graph import build_from_edges
    src = torch.arange(5)
    tgt = torch.zeros(5, dtype=torch.long)
    with pytest.raises(DgnnError):
        build_from_edges(torch.stack([src, tgt]), None, None, 5, 5, "cuda:0")


def test_empty_graph():
    from dgnn_b200.graph import build_from_edges
    eg = build_from_edges(torch.zeros((2, 0), dtype=torch.long), None, None, 3, 3, "cuda:0")
    assert (eg.nbr.cpu().numpy() == -1).all()


def test_labels_and_interface_facets_match_oracle(g):
    from dgnn_b200 import runModel as rm
    rng = np.random.default_rng(1)
    z = rng.standard_normal((g["n"], 2)).astype(np.float32)
    z[::7, 1] = z[::7, 0]                           # exact ties -> label 0
    lab = rm.labels(torch.from_numpy(z).cuda())
    assert np.array_equal(lab.cpu().numpy(), og.labels_from_logits(z))
    nfin = int((g["infinite"] == 0).sum())
    nf = rng.integers(-1, nfin, size=(5000, 2)).astype(np.int32)
    flag = rm.interface_facets(lab[:nfin].contiguous(), torch.from_numpy(nf))
    assert np.array_equal(np.nonzero(flag.cpu().numpy())[0], og.interface_facets(og.labels_from_logits(z)[:nfin], nf))


@pytest.mark.parametrize("od", [1, 2])
def test_export_scores_match_torch(od):
    """Row P, ``dataLoader.exportScore``: sigmoid / softmax of the logits as the reference writes them."""
    from dgnn_b200 import runModel as rm
    torch.manual_seed(0)
    z = torch.randn(5000, od) * 6
    out = rm.export_scores(z.to("cuda:0"))
    assert out["number_of_cells"] == 5000
    np.testing.assert_allclose(out["sigmoid"], z.sigmoid().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(out["softmax"], z.softmax(dim=-1).numpy(), rtol=2e-6, atol=1e-7)
    assert np.array_equal(out["logits"], z.numpy())
