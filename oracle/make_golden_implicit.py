"""TEST INFRASTRUCTURE — known answers of the reference's (dead) implicit-integration path, from the UNMODIFIED reference.

    make -C oracle ref && python oracle/make_golden_implicit.py        -> tests/golden/kat_implicit.npz

* Energy / ElasticPotential (material_point_method.cpp:160-209) on config 1's state after 200 substeps (2147 particles,
  impact under way: FE, FP far from identity) at several trial-velocity perturbations and time steps,
* the vendored optimiser (external/mcloptlib LBFGS + Backtracking, configured as mathy.hpp:10-19, convergence rule of
  mathy.hpp:31-35) on a smooth test function with an analytic gradient,
* timeIntegration (cpp:211-233) itself on a 24-particle subset: what the reference's finite-difference search does.
Needs /root/reference (build container only); the .npz is committed and is what travels."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "ref_mpm")
GOLD = os.path.join(HERE, "..", "tests", "golden")
SEED = 20260117


def run(args):
    r = subprocess.run([REF] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.exit("ref_mpm failed: " + r.stdout)
    return r.stdout


def main():
    c1 = np.load(os.path.join(GOLD, "c1_default.npz"))
    I = J = K = 20
    rng = np.random.default_rng(SEED)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- Energy ----
        state = np.ascontiguousarray(c1["state200"], np.float32)
        n = state.shape[0]
        state.tofile(tmp + "/state.f32")
        base = rng.standard_normal((I * J * K, 3)).astype(np.float32)
        cases = [(0.0, 1e-5), (0.05, 1e-5), (1.0, 1e-5), (0.05, 1e-3), (1.0, 1e-3), (2.0, 1e-3)]      # (amplitude [m/s], dt); beyond ~3 m/s x 1e-3 s trial states invert (det < 0), where the polar factor is convention-dependent
        res = []
        for amp, dt in cases:
            (base * np.float32(amp)).astype(np.float32).tofile(tmp + "/pert.f32")
            run(["--grid", I, J, K, "--n", n, "--dt", dt, "--load-full", tmp + "/state.f32", "--no-colliders", "--quiet",
                 "--kat-energy", tmp + "/pert.f32", tmp + "/e.f32"])
            res.append(np.fromfile(tmp + "/e.f32", dtype=np.float32))
        out.update(energy_state=state, energy_pert_unit=base, energy_cases=np.array(cases, np.float64), energy_ref=np.array(res, np.float32))
        print("Energy, ElasticPotential, used cells per case:\n", np.array(res))
        # ---- optimiser ----
        dim = 24
        a = rng.uniform(0.5, 4.0, dim).astype(np.float32)
        t = rng.uniform(-2.0, 2.0, dim).astype(np.float32)
        x0 = rng.uniform(-3.0, 3.0, dim).astype(np.float32)
        np.concatenate([[np.float32(dim)], a, t, x0]).astype(np.float32).tofile(tmp + "/lb.f32")
        run(["--kat-lbfgs", tmp + "/lb.f32", tmp + "/lbo.f32"])
        lb = np.fromfile(tmp + "/lbo.f32", dtype=np.float32)
        out.update(lbfgs_a=a, lbfgs_t=t, lbfgs_x0=x0, lbfgs_iters=np.int32(lb[0]), lbfgs_f=lb[1], lbfgs_x=lb[2:])
        print("LBFGS iterations", int(lb[0]), "f", lb[1])
        # ---- timeIntegration on a small subset, slow particles (|v| ~ 1 m/s so the library's eps = 2.2e-6 is representable) ----
        sub = state[rng.choice(n, 24, replace=False)].copy()
        sub[:, 1:4] *= np.float32(0.01)
        sub.tofile(tmp + "/sub.f32")
        run(["--grid", I, J, K, "--n", 24, "--dt", 1e-3, "--load-full", tmp + "/sub.f32", "--no-colliders", "--quiet",
             "--kat-implicit", tmp + "/gi.f32"])
        gi = np.fromfile(tmp + "/gi.f32", dtype=np.float32).reshape(-1, 7)
        out.update(ti_state=sub, ti_dt=np.float32(1e-3), ti_grid_after=gi[:, [0, 4, 5, 6]])
    np.savez_compressed(os.path.join(GOLD, "kat_implicit.npz"), **out)
    print("wrote kat_implicit.npz")


if __name__ == "__main__":
    main()
