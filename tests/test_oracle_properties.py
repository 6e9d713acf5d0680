"""CPU property tests of the oracle restatement (independent of the golden vectors): algebraic invariants of the
reference's algorithm that any faithful restatement must satisfy."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle_py as op
import mpm_b200
from scene_util import oracle_from_scene


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(-2.0, 2.0, allow_nan=False, width=32), min_size=9, max_size=9))
def test_svd3_reconstructs_and_orders(vals):
    A = np.array(vals, np.float32).reshape(3, 3)
    rc, U, S, V = op.svd3(A)
    assert rc == 0
    assert (np.diff(S) <= 0).all() and (S >= 0).all(), "Eigen sorts singular values in descending order"
    scale = max(np.abs(A).max(), 1e-6)
    assert np.abs(U @ np.diag(S) @ V.T - A).max() <= 2e-5 * scale + 1e-6
    assert np.abs(U.T @ U - np.eye(3)).max() < 5e-6 and np.abs(V.T @ V - np.eye(3)).max() < 5e-6
    assert np.abs(S - np.linalg.svd(A.astype(np.float64), compute_uv=False)).max() <= 2e-5 * scale + 1e-6


def test_svd3_rejects_non_finite_input_like_eigen():
    A = np.eye(3, dtype=np.float32)
    A[1, 2] = np.nan
    assert op.svd3(A)[0] == 1
    A[1, 2] = np.inf
    assert op.svd3(A)[0] == 1


@settings(max_examples=300, deadline=None)
@given(st.floats(0.0, 0.9999899864196777, allow_nan=False, width=32))
def test_weights_partition_of_unity(fx):
    # the four non-zero cubic B-spline weights of a particle at fractional offset fx sum to 1 (hpp:20-31)
    w = [op.lib().oracle_weight(float(np.float32(fx) - d)) for d in (-1, 0, 1, 2)]
    assert abs(sum(w) - 1.0) < 3e-7
    assert op.lib().oracle_weight(float(np.float32(fx) + 2)) == 0.0      # the fifth enumerated node (cpp:84-91) weighs nothing


def test_p2g_conserves_mass_and_momentum_and_g2p_reproduces_uniform_flow():
    sc = mpm_b200.scenes.small_ball(grid=32, radius_cells=4.0, with_ground=False)
    sc["vel"][:] = (3.0, -2.0, 5.0)
    o, cols, nc = oracle_from_scene(sc)
    g = o.grid()
    assert abs(g[:, 0].astype(np.float64).sum() - sc["mass"].astype(np.float64).sum()) < 1e-6 * sc["mass"].sum()
    mom = (g[:, 4:7].astype(np.float64) * g[:, 0:1]).sum(0)
    assert np.abs(mom - (sc["vel"].astype(np.float64) * sc["mass"][:, None]).sum(0)).max() < 1e-5 * np.abs(mom).max()
    o.g2p()
    s = o.state()
    assert np.abs(s[:, 1:4] - np.array([3.0, -2.0, 5.0], np.float32)).max() < 1e-4, "uniform field must be reproduced"
    assert np.abs(s[:, 26:35]).max() < 1e-5, "APIC matrix of a uniform field is zero up to cancellation"


def test_plasticity_clamp_bounds_and_det_split():
    rng = np.random.default_rng(7)
    n = 512
    st35 = np.zeros((n, 35), np.float32)
    st35[:, 0] = 6e-5; st35[:, 4] = 3e-5; st35[:, 5:8] = 0.5
    st35[:, 8] = st35[:, 12] = st35[:, 16] = 1.0
    st35[:, 17] = st35[:, 21] = st35[:, 25] = 1.0
    st35[:, 26:35] = rng.normal(scale=40.0, size=(n, 9)).astype(np.float32)      # large velocity gradients
    o = op.Oracle(20, 20, 20, n)
    o.set_state(st35)
    assert o.fupdate(1e-5) == 0
    s = o.state()
    FE = s[:, 8:17].reshape(-1, 3, 3).astype(np.float64); FP = s[:, 17:26].reshape(-1, 3, 3).astype(np.float64)
    sv = np.linalg.svd(FE, compute_uv=False)
    assert sv.max() <= 1.005 + 1e-5 and sv.min() >= 0.975 - 1e-5
    # det(FE * FP) equals det((I + dt B D^-1)) of the input: the split moves volume change between FE and FP only
    B = st35[:, 26:35].reshape(-1, 3, 3).astype(np.float64)       # glm column-major: transpose is irrelevant for det
    dinv = 1.0 / (0.05 * 0.05 / 3.0)
    want = np.linalg.det(np.eye(3) + 1e-5 * dinv * B)
    got = np.linalg.det(FE) * np.linalg.det(FP)
    assert np.abs(got - want).max() < 5e-5


def test_friction_quirk_is_reproduced():
    # cpp:290-291: vt.length() is the component count 3 -> vrel = vt * (1 + mu*vn/3) while 3 > -mu*vn, else 0
    w2l = np.eye(4, dtype=np.float32).reshape(1, 16)
    cols, nc = op.make_colliders(w2l, [[1.0, 1.0, 1.0]])
    import ctypes as C
    fp = C.POINTER(C.c_float)
    pos = np.array([0.0, 0.99, 0.0], np.float32)              # just inside the top face of a unit box at the origin
    for vy, expect_zero in ((-2.0, False), (-7.0, True)):
        vel = np.array([1.0, vy, 0.5], np.float32)
        out = np.zeros(3, np.float32)
        op.lib().oracle_body_collision(pos.ctypes.data_as(fp), vel.ctypes.data_as(fp), cols, nc, 0.5, out.ctypes.data_as(fp))
        if expect_zero:
            assert (out == 0).all()
        else:
            vn = vy                                            # normal ~ (0,1,0): vn = vy
            f = 1.0 + 0.5 * vn / 3.0
            assert abs(out[0] - 1.0 * f) < 1e-3 and abs(out[2] - 0.5 * f) < 1e-3 and abs(out[1]) < 1e-3


def test_quadratic_stencil_mode_of_the_oracle():
    """The oracle's quadratic mode (the checker for MpmParams.stencil = 1; not reference behaviour): partition of unity of the
    three-node weights, D = h^2/4, and mass / momentum conservation of one rasterisation + gather round trip."""
    import ctypes as C
    L = op.lib()
    L.oracle_weight_quadratic.restype = C.c_float
    L.oracle_weight_quadratic.argtypes = [C.c_float]
    for fx in np.linspace(-0.5, 0.499, 41):
        w = [L.oracle_weight_quadratic(float(np.float32(fx) - d)) for d in (-1, 0, 1)]
        assert abs(sum(w) - 1.0) < 2e-7 and min(w) >= 0.0
        assert L.oracle_weight_quadratic(float(np.float32(fx) - 2)) == 0.0 and L.oracle_weight_quadratic(float(np.float32(fx) + 2)) == 0.0
    rng = np.random.default_rng(5)
    n = 500
    pos = (rng.uniform(0.3, 0.7, (n, 3))).astype(np.float32)
    vel = rng.normal(0, 3, (n, 3)).astype(np.float32)
    o = op.Oracle(20, 20, 20, n, op.default_params(stencil=1))
    o.set_state(op.initial_state(pos, vel, np.float32(6e-5)))
    o.rasterize()
    g = o.grid().astype(np.float64)
    assert abs(g[:, 0].sum() - n * float(np.float32(6e-5))) < 1e-9
    mom = (g[:, 0:1] * g[:, 4:7]).sum(0)
    assert np.allclose(mom, (float(np.float32(6e-5)) * vel.astype(np.float64)).sum(0), rtol=1e-5, atol=1e-9)
    oc = op.Oracle(20, 20, 20, n, op.default_params())
    oc.set_state(op.initial_state(pos, vel, np.float32(6e-5)))
    oc.rasterize()
    assert (oc.grid()[:, 0] != 0).sum() > (g[:, 0] != 0).sum()          # fewer nodes touched than with the cubic stencil
