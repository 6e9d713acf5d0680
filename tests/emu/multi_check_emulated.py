"""TEST INFRASTRUCTURE ONLY: tools/multi_check.py without GPUs -- the slab-decomposed run over N processes (gloo, CPU
tensors, the host-emulation build of the library) against the same scene in one domain. Exercises the REAL slab protocol
of realtime-deformations_b200/multi.py and the library's halo / migration entry points on every rank count.
  MPM_B200_LIB=tests/emu/_build/libmpm_b200_emu.so MPM_B200_ALLOW_EMULATION=1 \
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/emu/multi_check_emulated.py [grid n steps balanced]"""
import os
import sys
from datetime import timedelta

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpm_b200                                    # noqa: E402
from importlib import import_module               # noqa: E402
multi = import_module("realtime-deformations_b200.multi")
from helpers import traj_errors                   # noqa: E402

grid, n, steps, balanced = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (64, 16384, 12, 0)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
assert os.environ.get("MPM_B200_ALLOW_EMULATION") == "1", "this script is for the host-emulation build only"
dist.init_process_group("gloo", timeout=timedelta(seconds=240))


def make_scene(i_range):
    sc = mpm_b200.scenes.snow_slab(grid=grid, n=n, i_range=i_range)
    sc["vel"][:] = (150.0, -20.0, 0.0)       # drive the slab across the slab boundaries
    return sc


n_layers = (grid + 3) // 4
if balanced == 2:
    # failure path: migration messages far too small for what crosses the slab boundaries -> the library flags the overflow,
    # the next collective count check must raise on EVERY rank in the same substep (nobody is left waiting in an exchange)
    multi.migrate_capacity_for = lambda n_local: 4
    lo, hi = multi.slab_layers(n_layers, world)[rank]
    r = multi.SlabRunner(grid, n, rank, world, torch, scene=make_scene((4 * lo + 1, 4 * hi + 1)), device="cpu")
    r.sync_every = 4
    caught_at = -1
    try:
        for k in range(steps):
            r.substep()
    except mpm_b200.capi.MpmError as exc:
        caught_at = r.steps_done
        print(f"rank {rank} raised at substep {caught_at}: {exc}", flush=True)
    got = [None] * world
    dist.all_gather_object(got, caught_at)
    if rank == 0:
        print("substep of the error per rank:", got)
        print("OVERFLOW_HANDLED" if (min(got) > 0 and len(set(got)) == 1) else "OVERFLOW_NOT_HANDLED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
if balanced:      # the bench's configuration: partition balanced by particle count, each rank generates its own cells
    r = multi.SlabRunner(grid, n, rank, world, torch, device="cpu")
else:
    lo, hi = multi.slab_layers(n_layers, world)[rank]
    r = multi.SlabRunner(grid, n, rank, world, torch, scene=make_scene((4 * lo + 1, 4 * hi + 1)), device="cpu")
r.sync_every = 4              # several collective count / overflow checks inside the run
if rank == 0:
    print("peer-memory halo:", bool(getattr(r, "peer_halo", False)))
n0 = r.sim.stats().n_particles
for _ in range(steps):
    r.substep()
st, pid = r.live_state()
gathered = [None] * world
dist.gather_object((st, pid, n0, st.shape[0], (r.lo, r.hi)), gathered if rank == 0 else None, dst=0)
if rank == 0:
    S = np.concatenate([g[0] for g in gathered]); P = np.concatenate([g[1] for g in gathered])
    print("layers per rank:", [g[4] for g in gathered])
    print("per-rank particles before/after:", [(g[2], g[3]) for g in gathered], "total", len(P), "unique ids", len(np.unique(P)))
    S = S[np.argsort(P)]
    one = multi.SlabRunner(grid, n, 0, 1, torch, scene=None if balanced else make_scene(None), device="cpu")
    for _ in range(steps):
        one.substep()
    ref = one.sim.download_state35()
    assert ref.shape[0] == S.shape[0], (ref.shape, S.shape)
    e = traj_errors(S, ref)
    print(f"multi({world}) vs single after {steps} substeps: max|dpos|={e[0]:.3e} max|dvel|={e[1]:.3e} max|ddetF|={e[2]:.3e}")
    alt = multi.SlabRunner(grid, n, 0, 1, torch, scene=None if balanced else make_scene(None), variants=(1, 1), device="cpu")
    for _ in range(steps):
        alt.substep()
    f = traj_errors(alt.sim.download_state35(), ref)
    print(f"noise floor (tile vs baseline kernels, one domain): max|dpos|={f[0]:.3e} max|dvel|={f[1]:.3e} max|ddetF|={f[2]:.3e}")
    vol_rel = np.abs(S[:, 4] / ref[:, 4] - 1).max()
    moved = sum(abs(g[2] - g[3]) for g in gathered)
    print("mass tag equal:", np.array_equal(S[:, 0], ref[:, 0]), " volumes max rel diff:", vol_rel, " migrated (net):", moved)
    ok = (len(np.unique(P)) == len(P) == ref.shape[0] and vol_rel < 1e-5 and (moved > 0 or balanced)
          and all(a <= 4 * max(b, c) for a, b, c in zip(e, f, (1e-6, 1e-3, 1e-5))))
    print("MULTI_CHECK_OK" if ok else "MULTI_CHECK_FAILED")
dist.barrier()
dist.destroy_process_group()
