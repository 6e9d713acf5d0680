"""TEST INFRASTRUCTURE: the plain-C oracle against the UNMODIFIED reference binary (oracle/_ref/ref_mpm, built from
/root/reference by oracle/Makefile) on randomized scenes -- beyond the committed golden vectors. Non-cubic grids, random
positions / velocities / masses / FE / FP / B, one to three rotated and moving box colliders, 1..6 substeps.
The oracle is bit-exact to the reference in every stage except the polar factor of the force stage (Newton in fp64 vs
Higham-Noferini in fp32, |dR| < 2e-6), so whole substeps agree to a few ulp; a case where that ulp flips the reference's
own elastic/plastic split is held to the total deformation instead.
    python oracle/fuzz_oracle_vs_ref.py [n_cases] [first_seed]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_py as op          # noqa: E402

REF = os.path.join(HERE, "_ref", "ref_mpm")


def total_F(s):
    return np.einsum("nij,njk->nik", s[:, 8:17].reshape(-1, 3, 3).transpose(0, 2, 1), s[:, 17:26].reshape(-1, 3, 3).transpose(0, 2, 1))


def run_case(seed, verbose=True):
    rng = np.random.default_rng(seed)
    h = float(rng.choice([0.05, 0.1]))
    I, J, K = (int(x) for x in rng.integers(9, 19, size=3))
    n = int(rng.integers(40, 500))
    lo, hi = 3.0 * h, np.array([(I - 3) * h, (J - 3) * h, (K - 3) * h])
    c = lo + rng.random(3) * (hi - lo)
    pos = np.clip(c + (rng.random((n, 3)) - 0.5) * 2 * h * rng.uniform(1.0, 4.0), lo, hi).astype(np.float32)
    pos[: n // 8] = (np.round(pos[: n // 8] / h) * h).astype(np.float32)          # some exactly on cell faces
    s0 = np.zeros((n, 35), np.float32)
    s0[:, 0] = (6e-5 * rng.uniform(0.5, 2.0, size=n)).astype(np.float32)
    s0[:, 1:4] = (rng.normal(size=(n, 3)) * rng.choice([1.0, 30.0])).astype(np.float32)
    s0[:, 5:8] = pos
    s0[:, 8:17] = (np.eye(3).reshape(-1) + rng.normal(size=(n, 9)) * 0.02).astype(np.float32)
    s0[:, 17:26] = (np.eye(3).reshape(-1) + rng.normal(size=(n, 9)) * 0.01).astype(np.float32)
    s0[:, 26:35] = (rng.normal(size=(n, 9)) * 1e-3).astype(np.float32)
    steps = int(rng.integers(1, 7))
    dt = 1e-5
    ncol = int(rng.integers(1, 4))
    rows = np.zeros((ncol, 7), np.float32)
    for k in range(ncol):
        rows[k, 0:3] = lo + rng.random(3) * (hi - lo)
        rows[k, 3] = rng.uniform(0.0, 90.0)
        rows[k, 4:7] = h * rng.uniform(1.0, 5.0, size=3)
    cvel = (rng.normal(size=3) * rng.choice([0.0, 3.0])).astype(np.float32)
    prm = op.default_params(h=h)
    # particle volumes: the oracle's own start-up stage (bit-exact to the reference's, tests/test_oracle_golden.py)
    o = op.Oracle(I, J, K, n, prm)
    o.set_state(s0)
    o.rasterize(); o.volumes()
    s0[:, 4] = o.state()[:, 4]
    with tempfile.TemporaryDirectory() as d:
        s0.tofile(os.path.join(d, "in.f32")); rows.tofile(os.path.join(d, "cols.f32"))
        cmd = [REF, "--grid", str(I), str(J), str(K), "--n", str(n), "--h", repr(h), "--dt", repr(dt), "--steps", str(steps),
               "--load-full", os.path.join(d, "in.f32"), "--colliders", os.path.join(d, "cols.f32"),
               "--collider-vel", repr(float(cvel[0])), repr(float(cvel[1])), repr(float(cvel[2])),
               "--dump-dir", d, "--dump-steps", str(steps), "--quiet"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        ref = np.fromfile(os.path.join(d, f"particles_step{steps:04d}.f32"), np.float32).reshape(n, 35)
        cols, nc = op.colliders_from_ref_dump(np.fromfile(os.path.join(d, "colliders.f32"), np.float32))
    o.set_state(s0)
    o.substep(dt, cols, nc, steps)
    got = o.state()
    fin = np.isfinite(ref).all(axis=1)
    assert np.isfinite(got[fin]).all(), f"seed {seed}: oracle non-finite where the reference is finite"
    a, b = got[fin], ref[fin]
    d_fe = np.abs(a[:, 8:17] - b[:, 8:17]).max(1)
    d_tot = np.abs(total_F(a) - total_F(b)).reshape(-1, 9).max(1)
    flipped = bool(((d_fe > 2e-5) & (d_tot < 2e-5 + 0.05 * d_fe)).any())
    vs = max(float(np.abs(b[:, 1:4]).max()), 1.0)
    e_pos, e_vel = float(np.abs(a[:, 5:8] - b[:, 5:8]).max()), float(np.abs(a[:, 1:4] - b[:, 1:4]).max())
    e_tot = float(d_tot.max()) if len(d_tot) else 0.0
    if flipped:
        assert e_pos < 1e-4 and e_vel < 0.5 and e_tot < 1e-3, f"seed {seed}: beyond a split flip: pos {e_pos:.2e} vel {e_vel:.2e} F {e_tot:.2e}"
    else:
        assert e_pos <= 2.5e-7 and e_vel <= 1e-4 * vs and float(d_fe.max()) <= 2e-5 and e_tot <= 2e-5, \
            f"seed {seed}: pos {e_pos:.2e} vel {e_vel:.2e} (scale {vs:.1f}) FE {float(d_fe.max()):.2e} F {e_tot:.2e}"
    exact = int(np.array_equal(a[:, 5:8], b[:, 5:8])) + int(np.array_equal(a[:, 1:4], b[:, 1:4]))
    if verbose:
        print(f"seed {seed:5d} ok: grid {I}x{J}x{K} h {h} n {n} colliders {nc} steps {steps}: |dpos| {e_pos:.1e} |dvel| {e_vel:.1e} |dFE| {float(d_fe.max()):.1e}"
              + (" [split flipped]" if flipped else "") + (" [pos+vel bit-exact]" if exact == 2 else ""), flush=True)
    o.close()
    return flipped


if __name__ == "__main__":
    if not os.path.exists(REF):
        raise SystemExit(f"{REF} is missing: make -C oracle ref (needs /root/reference)")
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    flips = sum(run_case(s) for s in range(first, first + n_cases))
    print(f"ORACLE_VS_REFERENCE_OK {n_cases} cases ({flips} with a flipped elastic/plastic split)")
