"""Trajectory error of the fused CUDA path against the oracle, in units of the scene's own noise floor (oracle vs oracle + FMA
contraction), for the tolerance-form and the bit-faithful F-update. python tools/fupdate_mode_errors.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import mpm_b200
from helpers import traj_errors
from scene_util import oracle_from_scene, sim_from_scene

th = min(os.cpu_count() or 1, 32)
cases = [("small ball 32^3", mpm_b200.scenes.small_ball(grid=32, radius_cells=5.0), (20, 60, 40)),
         ("contact 64^3 128Ki", mpm_b200.scenes.stiff_snowball(grid=64, n=1 << 17, dt=1e-5, gap_cells=0.25), (10, 20)),
         ("collision 64^3 128Ki", mpm_b200.scenes.snowball_collision(grid=64, n=1 << 17, gap_cells=0.4), (10, 20))]
for name, sc, steps in cases:
    o, oc, onc = oracle_from_scene(sc, threads=th)
    of, _, _ = oracle_from_scene(sc, fma=True, threads=th)
    sims = {m: sim_from_scene(sc, fupdate_exact=m) for m in (0, 1)}
    done = 0
    for n in steps:
        o.substep(float(sc["dt"]), oc, onc, n); of.substep(float(sc["dt"]), oc, onc, n)
        done += n
        floor = traj_errors(o.state(), of.state())
        for m, (sim, cols, nc) in sims.items():
            sim.substep(float(sc["dt"]), cols, nc, n)
            e = traj_errors(sim.download_state35(), o.state())
            print(f"{name:22s} step {done:4d} fupdate_exact={m}: err/floor pos {e[0] / max(floor[0], 1e-30):6.2f} vel {e[1] / max(floor[1], 1e-30):6.2f} detF {e[2] / max(floor[2], 1e-30):6.2f}"
                  f"   (floor {floor[0]:.2e} {floor[1]:.2e} {floor[2]:.2e})")
