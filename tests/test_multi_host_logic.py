"""CPU tests (gloo, world_size 2) of the host side of the slab decomposition: the partition and the neighbour exchange
helpers of realtime-deformations_b200/multi.py, which are backend-agnostic (they only see torch tensors)."""
import os
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mpm_b200  # noqa: F401  (puts the repo root on sys.path)

multi = import_module("realtime-deformations_b200.multi")


def test_slab_layers_partition():
    for n_layers in (5, 32, 128, 129):
        for world in (1, 2, 3, 4, 8):
            if world > n_layers:
                continue
            parts = multi.slab_layers(n_layers, world)
            assert parts[0][0] == 0 and parts[-1][1] == n_layers
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])), "slabs must tile the layers contiguously"
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def test_slab_scene_generation_is_a_partition_of_the_global_scene():
    full = mpm_b200.scenes.snow_slab(grid=64, n=100000)
    parts = multi.slab_layers(16, 3)
    pos = [mpm_b200.scenes.snow_slab(grid=64, n=100000, i_range=(4 * lo + 1, 4 * hi + 1))["pos"] for lo, hi in parts]
    assert np.array_equal(np.concatenate(pos), full["pos"]), "per-rank generation must reproduce the global particle list"
    cells = (np.concatenate(pos)[:, 0] / np.float32(0.05)).astype(np.int32)
    owner = np.concatenate([np.full(len(p), r) for r, p in enumerate(pos)])
    lo = np.array([p[0] for p in parts]); hi = np.array([p[1] for p in parts])
    layer = (cells - 1) >> 2
    assert ((layer >= lo[owner]) & (layer < hi[owner])).all(), "every particle starts on the rank that owns its block layer"


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # halo-style exchange: fixed-size buffers, ghost layer goes up, first layer goes down
        up = torch.full((8,), float(10 + rank)); dn = torch.full((8,), float(20 + rank))
        r_up, r_dn = torch.zeros(8), torch.zeros(8)
        multi.exchange_with_neighbours(dist, torch, rank, world, dn if rank > 0 else None, up if rank < world - 1 else None,
                                       r_dn if rank > 0 else None, r_up if rank < world - 1 else None)
        if rank < world - 1:
            assert (r_up == 20 + rank + 1).all()        # the upper neighbour's "down" buffer
        if rank > 0:
            assert (r_dn == 10 + rank - 1).all()        # the lower neighbour's "up" buffer
        # migration-style exchange: counts first, then ragged payloads (zero-length messages are skipped on both sides)
        n_dn, n_up = (0 if rank == 0 else 3 + rank), (0 if rank == world - 1 else 5 * rank)
        in_dn, in_up = multi.exchange_counts(dist, torch, rank, world, n_dn, n_up, "cpu")
        assert in_dn == (5 * (rank - 1) if rank > 0 else 0) and in_up == (3 + rank + 1 if rank < world - 1 else 0)
        F = 4
        s_dn = torch.arange(n_dn * F, dtype=torch.float32) + 100 * rank if n_dn else None
        s_up = torch.arange(n_up * F, dtype=torch.float32) + 1000 * rank if n_up else None
        g_dn = torch.zeros(in_dn * F) if in_dn else None
        g_up = torch.zeros(in_up * F) if in_up else None
        multi.exchange_with_neighbours(dist, torch, rank, world, s_dn, s_up, g_dn, g_up)
        if in_dn:
            assert torch.equal(g_dn, torch.arange(in_dn * F, dtype=torch.float32) + 1000 * (rank - 1))
        if in_up:
            assert torch.equal(g_up, torch.arange(in_up * F, dtype=torch.float32) + 100 * (rank + 1))
        out.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_balanced_partition_follows_particle_counts():
    counts = mpm_b200.scenes.snow_slab_layer_counts(512, 1 << 26)
    assert counts.sum() == 1 << 26
    for world in (2, 4, 8):
        parts = multi.slab_layers_balanced(counts, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(counts)
        assert all(a[1] == b[0] and a[1] > a[0] for a, b in zip(parts, parts[1:]))
        per_rank = np.array([counts[lo:hi].sum() for lo, hi in parts], np.float64)
        assert per_rank.max() / per_rank.mean() < 1.06, "particle counts per rank must be balanced to a few percent"
    # degenerate inputs: all particles in one layer, more ranks than occupied layers
    one = np.zeros(16, np.int64); one[5] = 1000
    parts = multi.slab_layers_balanced(one, 4)
    assert parts[0][0] == 0 and parts[-1][1] == 16 and all(hi > lo for lo, hi in parts)
