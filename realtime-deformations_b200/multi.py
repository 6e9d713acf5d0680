"""Slab decomposition of one scene over the GPUs of one box: one process per GPU (torch.distributed, NCCL over
NVLink / NVSwitch), the grid split along i into contiguous runs of 4-cell block layers.

Per substep each rank runs  bin/clear/P2G  ->  [halo exchange]  ->  grid update / F-update / G2P  ->  [migration].
  halo      : a rank's P2G also writes the first block layer of its upper neighbour (its "ghost" layer). The ghost
              layer's partial sums go up, the neighbour's own partial sums for that layer come down, both sides add
              (a+b == b+a in IEEE), and both then run the identical grid update on it: ONE exchange, no second trip.
              A layer is one contiguous chunk of the blocked grid (nbj*nbk*64 float4), 16.6 MB at 512^3.
  migration : particles whose block layer left the slab are packed by the library (44 floats each), counts and
              payloads go to the two neighbours with grouped isend/irecv, and are appended behind the live particles.
The comm helpers below only see torch tensors, so the same code runs under gloo on CPU tensors (tests).
"""
import numpy as np

from . import capi, scenes


def slab_layers(n_layers, world):
    """Contiguous, near-equal split of the particle-block layers along i: rank r owns [lo, hi)."""
    base, rem = divmod(n_layers, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def slab_layers_balanced(layer_counts, world):
    """Contiguous split of the block layers with near-equal PARTICLE counts (layer_counts[l] = particles whose block
    layer is l). Every rank gets at least one layer."""
    c = np.asarray(layer_counts, np.float64)
    n_layers = len(c)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target))
        # nearest layer boundary to the target, keeping at least one layer per remaining rank
        k = k if abs(cum[min(k, n_layers)] - target) <= abs(cum[max(k - 1, 0)] - target) else k - 1
        k = min(max(k, cuts[-1] + 1), n_layers - (world - r))
        cuts.append(k)
    cuts.append(n_layers)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def migrate_capacity_for(n_local):
    """Records per migration message wanted by a rank holding n_local particles. A plane of the 8-ppc slab sheds about
    8 * cells_in_plane * |v| dt / h particles per substep (hundreds at avalanche speeds, thousands at 200 m/s); n/512
    leaves an order of magnitude of head-room while the four fixed-size messages per substep stay at a few MB. The ranks
    then agree on the MAXIMUM of these (the message size must be identical on both ends of every exchange)."""
    return min(max(1 << 13, n_local // 512), 1 << 17)


def exchange_with_neighbours(dist, torch, rank, world, send_down, send_up, recv_down, recv_up):
    """Grouped point-to-point exchange with rank-1 (down) and rank+1 (up); any tensor may be None at the ends."""
    ops = []
    if rank > 0:
        if send_down is not None: ops.append(dist.P2POp(dist.isend, send_down, rank - 1))
        if recv_down is not None: ops.append(dist.P2POp(dist.irecv, recv_down, rank - 1))
    if rank < world - 1:
        if send_up is not None: ops.append(dist.P2POp(dist.isend, send_up, rank + 1))
        if recv_up is not None: ops.append(dist.P2POp(dist.irecv, recv_up, rank + 1))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def exchange_counts(dist, torch, rank, world, n_down, n_up, device):
    """Each rank tells its neighbours how many particles are coming; returns (incoming_from_down, incoming_from_up)."""
    sd = torch.tensor([n_down], dtype=torch.int64, device=device)
    su = torch.tensor([n_up], dtype=torch.int64, device=device)
    rd = torch.zeros(1, dtype=torch.int64, device=device)
    ru = torch.zeros(1, dtype=torch.int64, device=device)
    exchange_with_neighbours(dist, torch, rank, world, sd, su, rd, ru)
    return int(rd.item()) if rank > 0 else 0, int(ru.item()) if rank < world - 1 else 0


class SlabRunner:
    """Owns one rank's slab of the snow-slab scene (BASELINE config 5) and advances it substep by substep."""

    def __init__(self, grid, n_particles, rank, world, torch, scene=None, dt=1e-5, variants=(0, 0), sim_factory=None,
                 device="cuda"):
        """sim_factory / device exist for the CPU (gloo) tests of this protocol: a stand-in object with capi.Sim's
        pointer-based slab methods and device="cpu"; the product path always uses capi.Sim on "cuda"."""
        self.torch, self.rank, self.world, self.dt = torch, rank, world, float(dt)
        self.device = device
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
        n_layers = (grid + 3) // 4
        if scene is None and world > 1:
            # balance by particle count: the slab scene's per-layer counts are known in closed form
            self.lo, self.hi = slab_layers_balanced(scenes.snow_slab_layer_counts(grid, n_particles), world)[rank]
        else:
            self.lo, self.hi = slab_layers(n_layers, world)[rank]
        self.sync_every = 16
        # each rank generates only the cells of its own slab (counter-based RNG keyed by cell id)
        if scene is None:
            i_range = None if world == 1 else (max(4 * self.lo + 1, 0), 4 * self.hi + 1)   # cells c with (c-1)>>2 in [lo,hi)
            scene = scenes.snow_slab(grid=grid, n=n_particles, i_range=i_range)
        self.scene = scene
        n = scene["n"]
        p = capi.default_params(h=float(scene["h"]), p2g_variant=variants[0], g2p_variant=variants[1])
        if "gravity" in scene:
            p.gravity[:] = [float(x) for x in scene["gravity"]]
        self.migrates = world > 1
        cap = n if world == 1 else int(n * 1.5) + (1 << 16)
        make = sim_factory if sim_factory is not None else capi.Sim
        self.sim = make(grid, grid, grid, n, p, slab=None if world == 1 else (self.lo, self.hi), capacity=cap)
        # one stream for everything: the library's kernels and torch's NCCL calls are ordered against each other only
        # if the library stream IS torch's current stream while the collectives are enqueued
        self.stream = None
        if device == "cuda":
            self.stream = torch.cuda.Stream()
            self.sim.set_stream(self.stream.cuda_stream)
        with self._on_stream():
            self._setup(scene, n, rank, world, torch)

    def _on_stream(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def _setup(self, scene, n, rank, world, torch):
        if world > 1:
            # ids unique across ranks: exclusive prefix of the per-rank counts
            counts = torch.zeros(world, dtype=torch.int64, device=self.device)
            counts[rank] = n
            self.dist.all_reduce(counts)
            self.sim.set_pid_base(int(counts[:rank].sum().item()))
        self.sim.upload(scene["pos"], scene["vel"], scene["mass"])
        self.cols, self.nc = capi.make_colliders(scene["w2l"], scene["half"], scene["cvel"])
        self.h2d_bytes_per_step = 88 * self.nc + 4
        self.steps_done = 0
        if world > 1:
            # the migration messages have a fixed size: every rank must use the SAME record capacity
            cap = torch.tensor([migrate_capacity_for(n)], dtype=torch.int64, device=self.device)
            self.dist.all_reduce(cap, op=self.dist.ReduceOp.MAX)
            self.sim.set_migrate_capacity(int(cap.item()))
            mb = self.sim.migrate_buffer_bytes() // 4
            self.m_recv_dn = torch.empty(mb, dtype=torch.float32, device=self.device)
            self.m_recv_up = torch.empty(mb, dtype=torch.float32, device=self.device)
            hb = self.sim.halo_bytes() // 4
            self.h_send_up = torch.empty(hb, dtype=torch.float32, device=self.device)
            self.h_recv_up = torch.empty(hb, dtype=torch.float32, device=self.device)
            self.h_send_dn = torch.empty(hb, dtype=torch.float32, device=self.device)
            self.h_recv_dn = torch.empty(hb, dtype=torch.float32, device=self.device)
        # start-up of the reference (main.cpp:53-54): one P2G for the particle volumes; needs the halo as well
        self.sim.rasterizeParticlesToGrid()
        if world > 1:
            self._halo()
        self.sim.computeParticleVolumesAndDensities()
        # EXPERIMENTAL, opt-in (MPM_B200_PEER_HALO=1, not yet run on hardware): the per-substep ghost-layer reduction done by
        # P2G itself over NVLink (CUDA IPC mappings of the neighbours' grids, device-side flags) instead of halo messages
        import os
        self.peer_halo = world > 1 and os.environ.get("MPM_B200_PEER_HALO") == "1" and \
            (self.device == "cuda" or os.environ.get("MPM_B200_ALLOW_EMULATION") == "1")     # (tests/emu: shared-memory "IPC")
        if self.peer_halo:
            mine = (self.sim.peer_export(), self.hi - self.lo)
            everyone = [None] * world
            self.dist.all_gather_object(everyone, mine)
            lower = everyone[rank - 1] if rank > 0 else (None, 0)
            upper = everyone[rank + 1] if rank < world - 1 else (None, 0)
            self.sim.peer_connect(lower[0], lower[1], upper[0], upper[1])
            # the migration the same way: neighbours READ the packed buffers through their mappings (pull), no message
            mig = [None] * world
            self.dist.all_gather_object(mig, self.sim.peer_export_migration())          # (down buffer, up buffer) handles
            self.sim.peer_connect_migration(mig[rank - 1][1] if rank > 0 else None, mig[rank + 1][0] if rank < world - 1 else None)
            self.dist.barrier()            # every rank has mapped its neighbours before the first remote red

    # ghost-layer partial sums up, first-layer partial sums down, add on both sides
    def _halo(self):
        t, s, r, w = self.torch, self.sim, self.rank, self.world
        if r < w - 1:
            s.halo_pack(1, self.h_send_up.data_ptr())
        if r > 0:
            s.halo_pack(0, self.h_send_dn.data_ptr())
        exchange_with_neighbours(self.dist, t, r, w, self.h_send_dn if r > 0 else None, self.h_send_up if r < w - 1 else None,
                                 self.h_recv_dn if r > 0 else None, self.h_recv_up if r < w - 1 else None)
        if r < w - 1:
            s.halo_add(1, self.h_recv_up.data_ptr())
        if r > 0:
            s.halo_add(0, self.h_recv_dn.data_ptr())

    def _view(self, ptr, n_floats):
        """torch view over a buffer owned by the library (no copy); the library's buffers keep their addresses, so the
        views are built once (a few tens of microseconds each, every substep otherwise)."""
        cache = self.__dict__.setdefault("_views", {})
        key = (ptr, n_floats)
        if key not in cache:
            cache[key] = self._make_view(ptr, n_floats)
        return cache[key]

    def _make_view(self, ptr, n_floats):
        if self.device != "cuda":       # CPU stand-in of the tests: a plain host pointer
            import ctypes
            return self.torch.from_numpy(np.ctypeslib.as_array((ctypes.c_float * n_floats).from_address(ptr)))
        class _A:
            pass
        a = _A()
        a.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        return self.torch.as_tensor(a, device="cuda")

    def _migrate(self):
        """Fixed-size packed buffers (count in a device-side header): no host synchronisation per substep; the launch
        bound is re-tightened and overflow checked every `sync_every` substeps."""
        t, s, r, w = self.torch, self.sim, self.rank, self.world
        if getattr(self, "peer_halo", False):
            s.migrate_peer(0); s.migrate_peer(1)
            return self._after_migration()
        p_dn, p_up = s.migrate_pack()
        nf = s.migrate_buffer_bytes() // 4
        exchange_with_neighbours(self.dist, t, r, w, self._view(p_dn, nf) if r > 0 else None, self._view(p_up, nf) if r < w - 1 else None,
                                 self.m_recv_dn if r > 0 else None, self.m_recv_up if r < w - 1 else None)
        if r > 0:
            s.migrate_append_packed(self.m_recv_dn.data_ptr())
        if r < w - 1:
            s.migrate_append_packed(self.m_recv_up.data_ptr())
        self._after_migration()

    def _after_migration(self):
        t, s = self.torch, self.sim
        self.steps_done += 1
        if self.steps_done % self.sync_every == 0:
            # a capacity error on ONE rank must stop ALL ranks together (otherwise the others wait in the next exchange)
            err, msg = 0, ""
            try:
                s.sync_counts()
            except capi.MpmError as exc:
                err, msg = 1, str(exc)
            flag = t.tensor([err], dtype=t.int32, device=self.device)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX)
            if int(flag.item()):
                raise capi.MpmError(msg or "a neighbouring rank ran out of migration / slab capacity")

    def substep(self, host_colliders=False):
        if host_colliders:       # e2e arm: the per-frame host inputs are rebuilt and handed over every step
            self.cols, self.nc = capi.make_colliders(self.scene["w2l"], self.scene["half"], self.scene["cvel"])
        if self.world == 1:
            self.sim.substep(self.dt, self.cols, self.nc, 1)
            return
        with self._on_stream():
            if getattr(self, "peer_halo", False):
                for phase in (0, 1, 2):
                    self.sim.substep_begin_peer(self.dt, phase)
            else:
                self.sim.substep_begin(self.dt)
                self._halo()
            self.sim.substep_end(self.dt, self.cols, self.nc)
            self._migrate()

    def download_positions(self, pinned_xyzs):
        n = min(pinned_xyzs.shape[0], int(self.sim.stats().n_particles))
        self.sim.L.mpm_download_render_buffers(self.sim.h, n, pinned_xyzs.numpy().ctypes.data, None, 0.02)

    def live_state(self):
        cap = int(self.sim.stats().n_particles) + 16
        return self.sim.download_live(cap)
