"""Sharded scenes and the distributed halo-map build on the CPU: the analytic lattice shard against the
whole lattice graph, a shard cut out of an ordinary graph against the NumPy oracle's re-layout, the boundary-first
renumbering, and - world size 4 on gloo - the maps negotiated between ranks against the oracle's global-table maps,
followed by a halo exchange through them."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import graph as og
from oracle.static_model import to_attr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")
DIMS = (4, 8, 4)


def test_lattice_codes_round_trip_and_regularity():
    from dgnn_b200 import scene as sc
    n = 2 * DIMS[0] * DIMS[1] * DIMS[2]
    gid = torch.arange(n)
    x, y, z, s = sc.lattice_decode(gid, DIMS)
    assert torch.equal(sc.lattice_encode(x, y, z, s, DIMS), gid)
    nb = sc.lattice_neighbours(gid, DIMS)
    assert nb.min() >= 0 and nb.max() < n
    # 4-regular, symmetric with reverse slot k == k, no self loops, four distinct neighbours
    for k in range(4):
        assert torch.equal(nb[nb[:, k], k], gid)
    assert bool((nb != gid[:, None]).all())
    assert all(len(set(r)) == 4 for r in nb.tolist())
    # the same bonding as the host generator (different numbering)
    adj, _, _ = og.lattice_graph(*DIMS)
    deg = np.bincount(adj[:, 1], minlength=n)
    assert (deg == 4).all()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_lattice_shards_tile_the_global_scene(world):
    from dgnn_b200 import scene as sc
    g = sc.lattice_global(DIMS)
    n = g["x"].shape[0]
    nbr, rslot = og.ell_from_adjacency(g["edge_index"].t().numpy().astype(np.int32))
    ea_in_ref, ea_own_ref = og.relayout_edges(g["edge_attr"].numpy(), nbr, rslot)
    seen = 0
    for rank in range(world):
        sh = sc.lattice_scene(DIMS, rank, world, "cpu", need_backward=True, chunk=100)
        assert sh.n_global == n and sh.lo == seen
        seen = sh.hi
        assert np.array_equal(sh.nbr_gid.numpy(), nbr[sh.lo:sh.hi])
        assert torch.equal(sh.x, g["x"][sh.lo:sh.hi, 1:]) and torch.equal(sh.w, g["x"][sh.lo:sh.hi, 0])
        assert torch.equal(sh.y, g["y"][sh.lo:sh.hi])
        assert np.array_equal(sh.ea_in.numpy(), ea_in_ref[sh.lo:sh.hi])
        assert np.array_equal(sh.ea_own.numpy(), ea_own_ref[sh.lo:sh.hi])
    assert seen == n


def test_scene_from_global_matches_oracle_relayout():
    from dgnn_b200 import scene as sc
    g = np.load(GOLDEN)
    d = to_attr(dict(x=torch.from_numpy(g["x"]), edge_attr=torch.from_numpy(g["ea"]), y=torch.from_numpy(g["y"]),
                     edge_index=torch.from_numpy(g["adj"].T.astype(np.int64)).contiguous()))
    n = d.x.shape[0]
    nbr0, rslot = og.ell_from_adjacency(g["adj"])
    shards = [sc.scene_from_global(d, r, 3, "cpu", need_backward=True) for r in range(3)]
    perm = torch.cat([s.caller_ids for s in shards]).numpy()          # new -> old: the order the builder chose (RCM)
    assert sorted(perm.tolist()) == list(range(n))
    nbr = og.apply_perm_ell(nbr0, perm)
    ea_in_ref, ea_own_ref = og.relayout_edges(g["ea"], nbr0, rslot, perm)
    for s in shards:
        assert np.array_equal(s.nbr_gid.numpy(), nbr[s.lo:s.hi])
        assert np.array_equal(s.ea_in.numpy(), ea_in_ref[s.lo:s.hi])
        assert np.array_equal(s.ea_own.numpy(), ea_own_ref[s.lo:s.hi])
        assert np.array_equal(s.x.numpy(), g["x"][perm[s.lo:s.hi], 1:])
        assert np.array_equal(s.w.numpy(), g["x"][perm[s.lo:s.hi], 0])


@pytest.mark.parametrize("world", [2, 8])
def test_boundary_first_renumbering(world):
    from dgnn_b200.partition import boundary_first, build_halo_maps, partition_bounds
    from dgnn_b200 import scene as sc
    n = 2 * DIMS[0] * DIMS[1] * DIMS[2]
    nb = sc.lattice_neighbours(torch.arange(n), DIMS).to(torch.int32)
    bounds = partition_bounds(n, world)
    for rank in range(world):
        m = build_halo_maps(nb, bounds, rank)
        b = boundary_first(m)
        assert 0 < b.n_boundary <= m.n_own and torch.equal(torch.sort(b.order).values, torch.arange(m.n_own))
        # rows that peers need are exactly the first n_boundary rows
        assert set(b.send_idx.tolist()) == set(range(b.n_boundary))
        # same graph: global id of every (row, slot) is unchanged
        rows_old = torch.cat([torch.arange(m.lo, m.hi), m.halo_gid])
        rows_new = torch.cat([rows_old[:m.n_own][b.order], m.halo_gid])
        assert torch.equal(rows_new[b.local_nbr.long()], rows_old[m.local_nbr.long()][b.order])
        # the send lists name the same global cells, peer by peer, in the same order
        assert torch.equal(rows_new[b.send_idx.long()], rows_old[m.send_idx.long()])
        # Morton order kept inside each class
        assert bool((b.order[:b.n_boundary][1:] > b.order[:b.n_boundary][:-1]).all())
        assert bool((b.order[b.n_boundary:][1:] > b.order[b.n_boundary:][:-1]).all())


def _worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgnn_b200 import scene as sc
        from dgnn_b200.partition import HaloComm, boundary_first, build_halo_maps_sharded, partition_bounds
        n = 2 * DIMS[0] * DIMS[1] * DIMS[2]
        sh = sc.lattice_scene(DIMS, rank, world, "cpu")
        bounds = partition_bounds(n, world)
        m = build_halo_maps_sharded(sh.nbr_gid, sh.lo, sh.hi, bounds, rank)
        nb_global = sc.lattice_neighbours(torch.arange(n), DIMS).numpy()
        ref = og.halo_maps(nb_global, bounds.numpy(), rank)
        assert np.array_equal(m.local_nbr.numpy(), ref["local_nbr"])
        assert np.array_equal(m.halo_gid.numpy(), ref["halo_gid"])
        assert m.recv_counts == ref["recv_counts"].tolist()
        off = 0
        for q in range(world):
            assert np.array_equal(m.send_idx[off:off + m.send_counts[q]].numpy(), ref["send_idx"][q])
            off += m.send_counts[q]
        # exchange through the boundary-first maps: halo rows receive their owners' rows
        b = boundary_first(m)
        gid_rows = torch.cat([torch.arange(m.lo, m.hi)[b.order], m.halo_gid]).to(torch.float32)
        h = torch.full((m.n_own + m.n_halo, 8), -1.0)
        h[:m.n_own] = gid_rows[:m.n_own, None] * 2.0 + torch.arange(8.0)[None, :]
        comm = HaloComm(b, n)
        comm.start(h); comm.finish()
        assert torch.equal(h[m.n_own:], gid_rows[m.n_own:, None] * 2.0 + torch.arange(8.0)[None, :])
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharded_halo_maps_world4_gloo():
    world = 4
    ok = mp.get_context("spawn").Array("i", [0] * world)
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
