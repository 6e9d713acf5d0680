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


# ---- the whole per-substep protocol of SlabRunner on 8 gloo ranks, against a CPU stand-in of the library -------------
class _StandInSim:
    """Same pointer-based slab methods as capi.Sim, backed by numpy. Every message carries (sender rank, step, kind) so
    that the receiver can check WHO it came from and WHEN; the migration buffer size depends on the capacity the ranks
    agreed on, exactly like the real library (unequal sizes between neighbours is the bug this test exists for)."""
    HALO_FLOATS = 4096

    def __init__(self, I, J, K, n, params, slab=None, capacity=None):
        import ctypes
        self.ct = ctypes
        self.n, self.slab, self.capacity = n, slab, capacity
        self.rank = None
        self.step = 0            # substeps begun
        self.halo_round = 0      # halo exchanges seen (1 at start-up + 1 per substep)
        self.mig_round = 0
        self.mcap = None
        self.errors = []
        self.out = {}

    def _arr(self, ptr, n):
        return np.ctypeslib.as_array((self.ct.c_float * n).from_address(ptr))

    # set-up
    def set_pid_base(self, base): self.pid_base = base
    def upload(self, pos, vel, mass): assert len(pos) == self.n
    def set_migrate_capacity(self, cap): self.mcap = int(cap)
    def migrate_buffer_bytes(self): return 16 * (1 + 11 * self.mcap)
    def halo_bytes(self): return 4 * self.HALO_FLOATS
    def rasterizeParticlesToGrid(self): pass
    def computeParticleVolumesAndDensities(self): pass
    def stats(self): raise AssertionError("no per-step host read-back expected")

    # halo: upper=1 is my ghost layer (goes to rank+1), upper=0 my first layer (goes to rank-1)
    def halo_pack(self, upper, ptr):
        self._arr(ptr, self.HALO_FLOATS)[:] = 1000 * self.rank + 10 * self.halo_round + upper
    def halo_add(self, upper, ptr):
        got = self._arr(ptr, self.HALO_FLOATS)
        sender = self.rank + 1 if upper else self.rank - 1
        want = 1000 * sender + 10 * self.halo_round + (0 if upper else 1)      # the neighbour's OTHER layer
        if not (got == want).all():
            self.errors.append(f"halo round {self.halo_round}: rank {self.rank} upper={upper} got {got[0]} want {want}")
        if upper == 0 or self.slab[1] == self.top:          # last add of this exchange
            pass

    def substep_begin(self, dt):
        self.step += 1
        self.halo_round += 1
    def substep_end(self, dt, cols, nc): pass

    def migrate_pack(self):
        nf = self.migrate_buffer_bytes() // 4
        self.out = {d: np.full(nf, float(1000 * self.rank + d), np.float32) for d in (0, 1)}
        for d in (0, 1):
            self.out[d][:1].view(np.int32)[0] = (7 * self.rank + self.step + d) % 5       # header: record count
        return self.out[0].ctypes.data, self.out[1].ctypes.data
    def migrate_append_packed(self, ptr):
        nf = self.migrate_buffer_bytes() // 4
        got = self._arr(ptr, nf)
        self.appends = getattr(self, "appends", [])
        self.appends.append((int(got[:1].view(np.int32)[0]), float(got[-1])))
    def sync_counts(self): self.syncs = getattr(self, "syncs", 0) + 1


def _protocol_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sims = []

        def factory(*a, **k):
            s = _StandInSim(*a, **k); s.rank = rank; s.top = None; sims.append(s); return s
        multi.migrate_capacity_for = lambda n_local: 16 + n_local // 64       # per-rank wishes now really differ
        r = multi.SlabRunner(64, 60000, rank, world, torch, sim_factory=factory, device="cpu")
        sim = sims[0]
        # balanced slabs: unequal particle counts, yet every rank must have agreed on one migration capacity
        caps = [None] * world
        dist.all_gather_object(caps, (sim.n, sim.mcap, sim.slab))
        assert len({c[1] for c in caps}) == 1, f"migration capacity differs across ranks: {caps}"
        assert caps[0][1] == max(16 + c[0] // 64 for c in caps), "the agreed capacity is the largest wish"
        assert len({c[0] for c in caps}) > 1, "the test needs unequal per-rank particle counts"
        assert [c[2][0] for c in caps[1:]] == [c[2][1] for c in caps[:-1]], "slabs must be contiguous"
        steps = 20
        for _ in range(steps):
            r.substep()
            # migration: what arrived must be the neighbours' buffers of THIS step
            want = []
            if rank > 0:
                want.append(((7 * (rank - 1) + sim.step + 1) % 5, float(1000 * (rank - 1) + 1)))     # lower neighbour's "up" buffer
            if rank < world - 1:
                want.append(((7 * (rank + 1) + sim.step + 0) % 5, float(1000 * (rank + 1) + 0)))     # upper neighbour's "down" buffer
            assert sim.appends == want, (rank, sim.step, sim.appends, want)
            sim.appends = []
        assert sim.errors == [], sim.errors
        assert getattr(sim, "syncs", 0) == steps // r.sync_every
        out.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [3, 8])
def test_slab_protocol_on_many_ranks_with_stand_in_library(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_protocol_worker, args=(r, world, 29650 + world, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), [r for r in res if r[1] != "ok"]
