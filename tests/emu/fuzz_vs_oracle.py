"""TEST INFRASTRUCTURE ONLY: randomized scenes through the host-emulation build of the library (the device code, compiled
for the host) against the CPU oracle. Non-cubic grids whose dimensions are not multiples of the 4-cell blocks, particles
clustered, on cell faces, at the position clamp and outside the grid, random velocities / deformation gradients, moving
colliders, every kernel variant. Usage (needs MPM_B200_LIB=<emulated build> MPM_B200_ALLOW_EMULATION=1):
    python tests/emu/fuzz_vs_oracle.py [n_cases] [first_seed]
Prints one line per case and exits non-zero on the first disagreement beyond the scene's own noise floor."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mpm_b200                                                       # noqa: E402
import oracle_py as op                                                # noqa: E402
from helpers import assert_traj_close_calibrated                      # noqa: E402
from scene_util import oracle_from_scene                              # noqa: E402

VARIANTS = [(0, 0), (1, 1), (2, 0), (1, 0), (0, 1)]


def random_scene(rng):
    h = float(rng.choice([0.05, 0.05, 0.025, 0.1]))
    dims = tuple(int(x) for x in rng.integers(9, 30, size=3))
    I, J, K = dims
    lo, hi = 3.0 * h, np.array([(I - 3) * h, (J - 3) * h, (K - 3) * h])
    n_clusters = int(rng.integers(1, 4))
    parts = []
    for _ in range(n_clusters):
        c = lo + rng.random(3) * (hi - lo)
        r = h * rng.uniform(0.8, 4.0)
        m = int(rng.integers(1, 900))
        parts.append(c + (rng.random((m, 3)) - 0.5) * 2 * r)
    pos = np.concatenate(parts)
    # special positions: exactly on cell faces, exactly at the clamp bounds, a crowded cell
    k = min(len(pos), 40)
    pos[:k] = np.round(pos[:k] / h) * h
    pos[k:k + 5, 0] = lo
    pos[k + 5:k + 10, 1] = hi[1]
    crowd = lo + rng.random(3) * (hi - lo)
    pos = np.concatenate([pos, crowd + rng.random((int(rng.integers(0, 700)), 3)) * h * 0.9])
    pos = np.clip(pos, lo, hi).astype(np.float32)
    n_out = int(rng.integers(1, 4)) if rng.random() < 0.25 else 0     # sometimes: particles outside the cells the reference can index
    if n_out:
        pos[-n_out:] = (rng.random((n_out, 3)) * h * 1.5).astype(np.float32)
    n = len(pos)
    vel = (rng.normal(size=(n, 3)) * rng.choice([0.5, 20.0, 150.0])).astype(np.float32)
    cols = []
    if rng.random() < 0.8:
        top = float(lo + rng.random() * (hi[1] - lo) * 0.5) + h / 2
        w2l, half, _ = mpm_b200.scenes.ground_collider(top, dims, h)
        cols.append((w2l, half, (rng.normal(size=3) * rng.choice([0.0, 3.0])).astype(np.float32)))
    if rng.random() < 0.4:                              # a small rotated box somewhere in the domain
        a = rng.uniform(0, np.pi)
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        t = lo + rng.random(3) * (hi - lo)
        M = np.eye(4); M[:3, :3] = R.T; M[:3, 3] = -R.T @ t          # inverse(translate(t) * rot)
        cols.append((M.T.astype(np.float32).reshape(16), (h * rng.uniform(1, 4, size=3)).astype(np.float32),
                     (rng.normal(size=3) * 2.0).astype(np.float32)))
    sc = mpm_b200.scenes._scene(pos, vel, dims, h, 1e-5 if h >= 0.05 else 5e-6, cols, name="fuzz")
    sc["mass"] = (sc["mass"] * rng.uniform(0.5, 2.0, size=n)).astype(np.float32)
    sc["n_out"] = n_out
    return sc


def run_case(seed):
    rng = np.random.default_rng(seed)
    sc = random_scene(rng)
    variants = VARIANTS[seed % len(VARIANTS)]
    steps = int(rng.integers(1, 9))
    if sc["n_out"]:
        # the reference indexes out of bounds for such particles (undefined); the library parks them for good, the oracle
        # skips them in the transfers but still clamps their position into the domain, where they interact from the
        # second substep on -- only the first substep is comparable
        steps = 1
    prm = dict(theta_c=float(rng.choice([2.5e-2, 1e-2])), theta_s=float(rng.choice([5e-3, 2e-3])), hardening_xi=float(rng.choice([10.0, 5.0])))
    p = mpm_b200.capi.default_params(h=float(sc["h"]), p2g_variant=variants[0], g2p_variant=variants[1], **prm)
    I, J, K = sc["dims"]
    sim = mpm_b200.Sim(I, J, K, sc["n"], p)
    sim.upload(sc["pos"], sc["vel"], sc["mass"])
    sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()
    cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
    oprm = dict(theta_c=prm["theta_c"], theta_s=prm["theta_s"], xi=prm["hardening_xi"])      # the oracle's field is called xi
    o, ocols, onc = oracle_from_scene(sc, **oprm)
    of, _, _ = oracle_from_scene(sc, fma=True, **oprm)
    # deformed start: random FE / FP near the identity go in through the full-state upload on both sides
    s0 = o.state().copy()
    s0[:, 8:17] += (rng.normal(size=(sc["n"], 9)) * 0.02).astype(np.float32)
    s0[:, 17:26] += (rng.normal(size=(sc["n"], 9)) * 0.01).astype(np.float32)
    g = sim.download_state35()
    assert np.array_equal(g[:, 5:8], s0[:, 5:8]) and np.array_equal(g[:, 0], s0[:, 0])
    vol_ok = np.abs(g[:, 4] - s0[:, 4]) <= 2e-5 * np.abs(s0[:, 4]).max()
    assert vol_ok.all(), f"volumes differ: {np.abs(g[:, 4] - s0[:, 4]).max()}"
    s0[:, 4] = g[:, 4]
    for x in (o, of):
        x.set_state(s0)
    sim.upload_state35(s0)
    dt = float(sc["dt"])
    sim.substep(dt, cols, nc, steps)
    o.substep(dt, ocols, onc, steps); of.substep(dt, ocols, onc, steps)
    a, b, c = sim.download_state35(), o.state(), of.state()
    live = slice(0, sc["n"] - sc["n_out"]) if sc["n_out"] else slice(None)
    fin = np.isfinite(b[live]).all(axis=1) & np.isfinite(c[live]).all(axis=1)
    assert np.isfinite(a[live][fin]).all(), "non-finite state where the oracle is finite"
    # The reference re-assembles FE from transposed SVD factors: for nearly equal singular values fp32
    # noise decides on which axis a value lands, FE and FP jump by O(1e-2) while FE*FP stays put, and the stress follows.
    # Such a flip between library and oracle is the reference's chaos, not a kernel error: the case is then only held to
    # the total deformation and a loose bound.
    def total_F(s):
        return np.einsum("nij,njk->nik", s[:, 8:17].reshape(-1, 3, 3).transpose(0, 2, 1), s[:, 17:26].reshape(-1, 3, 3).transpose(0, 2, 1))
    A, Bq = a[live][fin], b[live][fin]
    d_fe = np.abs(A[:, 8:17] - Bq[:, 8:17]).max(1)
    d_tot = np.abs(total_F(A) - total_F(Bq)).reshape(-1, 9).max(1)
    flipped = bool(((d_fe > 1e-4) & (d_tot < 2e-5 + 0.05 * d_fe)).any())
    if flipped:
        assert np.abs(A[:, 5:8] - Bq[:, 5:8]).max() < 1e-4 and np.abs(A[:, 1:4] - Bq[:, 1:4]).max() < 0.5, f"fuzz seed {seed}: beyond a split flip"
    else:
        assert_traj_close_calibrated(A, Bq, c[live][fin], f"fuzz seed {seed}")
    st = sim.stats()
    assert st.n_particles == sc["n"] and st.n_out_of_grid == o.num_out_of_grid(), (st.n_particles, st.n_out_of_grid, o.num_out_of_grid())
    if sc["n_out"]:
        assert np.array_equal(a[-sc["n_out"]:, 5:8], s0[-sc["n_out"]:, 5:8]), "parked particles must not move"
    print(f"seed {seed:5d} ok: dims {sc['dims']} h {float(sc['h']):.3f} n {sc['n']:5d} (+{sc['n_out']} parked) colliders {nc} variants {variants} steps {steps}"
          + (" [elastic/plastic split flipped by fp32 noise: loose bound]" if flipped else ""), flush=True)
    sim.close()


if __name__ == "__main__":
    assert os.environ.get("MPM_B200_ALLOW_EMULATION") == "1", "point MPM_B200_LIB at the host-emulation build"
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for seed in range(first, first + n_cases):
        run_case(seed)
    print(f"FUZZ_OK {n_cases} cases")
