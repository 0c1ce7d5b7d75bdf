"""Partition / halo maps against the NumPy oracle, and the halo exchange with world_size 2 on gloo (CPU)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import graph as og

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")


def _global_table():
    g = np.load(GOLDEN)
    nbr, _ = og.ell_from_adjacency(g["adj"])
    perm = og.morton_perm(g["cen"])
    return og.apply_perm_ell(nbr, perm)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_halo_maps_bit_exact_vs_oracle(world):
    from dgnn_b200.partition import build_halo_maps, partition_bounds
    nb = _global_table()
    n = nb.shape[0]
    bounds = partition_bounds(n, world)
    assert np.array_equal(bounds.numpy(), og.partition_ranges(n, world))
    for rank in range(world):
        m = build_halo_maps(torch.from_numpy(nb), bounds, rank)
        ref = og.halo_maps(nb, bounds.numpy(), rank)
        assert np.array_equal(m.local_nbr.numpy(), ref["local_nbr"])
        assert np.array_equal(m.halo_gid.numpy(), ref["halo_gid"])
        assert m.recv_counts == ref["recv_counts"].tolist()
        off = 0
        for q in range(world):
            cnt = m.send_counts[q]
            assert np.array_equal(m.send_idx[off:off + cnt].numpy(), ref["send_idx"][q])
            off += cnt


def _worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgnn_b200.partition import build_halo_maps, exchange_halo, partition_bounds
        nb = _global_table()
        n = nb.shape[0]
        bounds = partition_bounds(n, world)
        m = build_halo_maps(torch.from_numpy(nb), bounds, rank)
        f = 12
        gid = torch.arange(n, dtype=torch.float32)
        h_global = gid[:, None] * 3.0 + torch.arange(f, dtype=torch.float32)[None, :]   # row content identifies the cell
        h = torch.full((m.n_own + m.n_halo, f), -1.0)
        h[:m.n_own] = h_global[m.lo:m.hi]
        exchange_halo(h, m)
        assert torch.equal(h[m.n_own:], h_global[m.halo_gid])
        # the local table resolves to the same rows the global table names
        rows = torch.cat([torch.arange(m.lo, m.hi), m.halo_gid])
        assert torch.equal(rows[m.local_nbr.long()], torch.from_numpy(nb[m.lo:m.hi]).long())
        # reverse exchange (gradients accumulated on the source side): every rank scatters 1 per (target, slot) into
        # its [owned | halo] rows; after reverse_add each owned cell holds its global in-degree as a source
        from dgnn_b200.partition import HaloComm
        d = torch.zeros((m.n_own + m.n_halo, 4))
        d.index_add_(0, m.local_nbr.reshape(-1).long(), torch.ones((m.n_own * 4, 4)))
        HaloComm(m, n).reverse_add(d)
        deg = np.bincount(nb.reshape(-1), minlength=n).astype(np.float32)     # times a cell is named as a neighbour
        assert np.array_equal(d[:m.n_own, 0].numpy(), deg[m.lo:m.hi])
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
