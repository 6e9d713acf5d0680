"""GPU parity tests of the implicit (optimisation-based) time integration, SURVEY 8 row f4: the CUDA path through the C ABI
(mpm_energy, mpm_energy_gradient, mpm_time_integration) against (i) the UNMODIFIED reference's Energy values
(tests/golden/kat_implicit.npz, material_point_method.cpp:160-209), (ii) the oracle's analytic gradient, (iii) the oracle's
restatement of the vendored optimiser on the same objective. Tolerances are written next to each comparison."""
import os

import numpy as np
import pytest

import mpm_b200
import oracle_py as op
from conftest import GOLDEN
from scene_util import oracle_from_scene, sim_from_scene, sim_from_state35

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kat():
    return dict(np.load(os.path.join(GOLDEN, "kat_implicit.npz")))


def _oracle_on(state, dims=(20, 20, 20)):
    o = op.Oracle(dims[0], dims[1], dims[2], state.shape[0])
    o.set_state(state)
    o.rasterize()
    return o


FORMS = [(0, 0), (1, 1)]      # block-tile form (TMA-staged gather, register-accumulated scatter) and the thread-per-particle baseline


@pytest.mark.parametrize("variants", FORMS)
def test_energy_matches_reference(kat, variants):
    # Energy and ElasticPotential of the unmodified reference at six trial fields; the device sums in double, the
    # reference in float (2147-term float sums: ~1e-6 relative)
    sim = sim_from_state35(kat["energy_state"], (20, 20, 20), variants=variants)
    sim.rasterizeParticlesToGrid()
    assert sim.stats().n_active_nodes == int(kat["energy_ref"][0, 2])          # used_cells.size()
    for (amp, dt), ref in zip(kat["energy_cases"], kat["energy_ref"]):
        pert = (kat["energy_pert_unit"] * np.float32(amp)).astype(np.float32)
        e, el = sim.energy(float(dt), pert, relative=True)
        assert abs(e - ref[0]) <= 5e-6 * abs(ref[0]), (amp, dt, e, ref[0])
        assert abs(el - ref[1]) <= 5e-6 * abs(ref[1]), (amp, dt, el, ref[1])
    e0, _ = sim.energy(1e-5)                                                   # trial = the grid's own velocities: no inertia term
    assert abs(e0 - kat["energy_ref"][0, 0]) <= 5e-6 * kat["energy_ref"][0, 0]


@pytest.mark.parametrize("variants", FORMS)
def test_energy_gradient_matches_oracle(kat, variants):
    st = kat["energy_state"]
    o = _oracle_on(st)
    used = o.used_cells()
    sim = sim_from_state35(st, (20, 20, 20), variants=variants)
    sim.rasterizeParticlesToGrid()
    for amp, dt in ((0.5, 1e-3), (0.05, 1e-5), (0.0, 1e-3)):
        pert = (kat["energy_pert_unit"] * np.float32(amp)).astype(np.float32)
        g = sim.energy_gradient(dt, pert, relative=True)
        ref = o.energy_gradient(o.grid()[used][:, 4:7] + pert[used], dt)
        scale = np.abs(ref).max()
        assert scale > 0
        assert np.abs(g[used] - ref).max() <= 2e-5 * scale, (amp, dt, np.abs(g[used] - ref).max(), scale)     # float vector reds of ~100 terms per node
        assert np.abs(np.delete(g, used, axis=0)).max() == 0.0                 # nodes without mass never move (used_cells, cpp:105-110)
        assert abs(float(g.sum(0) @ np.ones(3)) - float(ref.sum())) <= 1e-4 * scale * np.sqrt(used.size)


@pytest.mark.parametrize("variants", FORMS)
def test_time_integration_matches_oracle_minimiser(kat, variants):
    # timeIntegration (cpp:211-233) on 24 slow particles: same optimiser, same objective, analytic gradient on both sides.
    # The iterates are sensitive (the second L-BFGS step scales by s.y / y.y of a 1e-4 first step), so the minimiser is held
    # to the energy it reaches (1 %) and to 2 % of the distance moved, not bit-wise.
    sub, dt = kat["ti_state"], float(kat["ti_dt"])
    o = _oracle_on(sub)
    used = o.used_cells()
    v_star = o.grid()[used][:, 4:7].copy()
    it_o, _ = o.time_integration(dt)
    v_o = o.grid()[used][:, 4:7]
    check = _oracle_on(sub)
    e0, e_o = check.energy(v_star, dt), check.energy(v_o, dt)

    sim = sim_from_state35(sub, (20, 20, 20), variants=variants)
    sim.rasterizeParticlesToGrid()
    st = sim.timeIntegration(dt)
    v_g = sim.grid()[used][:, 4:7]
    e_g = check.energy(v_g, dt)
    assert abs(st.iterations - it_o) <= 1 and st.iterations > 0, (st.iterations, it_o)
    assert abs(st.energy_start - e0) <= 1e-5 * e0 and abs(st.energy_end - e_g) <= 1e-5 * e_g
    assert e_g < 0.25 * e0 and abs(e_g - e_o) <= 1e-2 * e_o, (e0, e_o, e_g)
    moved = np.abs(v_o - v_star).max()
    assert np.abs(v_g - v_o).max() <= 2e-2 * moved, (np.abs(v_g - v_o).max(), moved)
    # the reference's own result (finite-difference search, golden) is no better than where it started
    assert check.energy(kat["ti_grid_after"][used][:, 1:4], dt) > 0.5 * e0
    # velocities of nodes without mass are untouched
    others = np.setdiff1d(np.arange(8000), used)
    assert np.abs(sim.grid()[others][:, 4:7]).max() == 0.0


def test_semi_implicit_substeps_on_a_ball_vs_oracle():
    # a 4 k-particle snowball mid-impact (70 explicit substeps on the GPU, state handed to both sides), then two substeps of:
    # rasterize | gravity | implicit elastic solve with the material of the explicit path (E = 1.4e5, nu = 0.2, hardening as
    # cpp:240) | collisions | F-update | G2P | advect, at a time step 5 x the explicit one. GPU and oracle run the same sequence.
    sc = mpm_b200.scenes.small_ball(grid=32, radius_cells=5.0)
    pre, cols, nc = sim_from_scene(sc)
    pre.substep(float(sc["dt"]), cols, nc, 70)
    state = pre.download_state35()
    assert np.abs(state[:, 8:17] - np.eye(3, dtype=np.float32).reshape(9)).max() > 1e-2        # really deformed
    dt = 5e-5
    E, nu = 1.4e5, 0.2
    mu0, lambda0 = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    o = op.Oracle(32, 32, 32, state.shape[0]); o.set_state(state)
    ocols, onc = op.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
    sim = sim_from_state35(state, (32, 32, 32))
    qo = op.default_implicit_params(mu0=mu0, lambda0=lambda0, hardening=1)
    qg = mpm_b200.capi.default_implicit_params(mu0=mu0, lambda0=lambda0, hardening=1)
    total_iters = 0
    for step in range(2):
        o.rasterize(); o.grid_velocities(dt)
        sim.rasterizeParticlesToGrid(); sim.gridVelocitiesUpdate(dt)
        g_before = sim.grid()
        go_before = o.grid()
        it_o, _ = o.time_integration(dt, qo)
        st = sim.timeIntegration(dt, qg)
        g_after = sim.grid()
        total_iters += st.iterations
        assert st.iterations > 0 and st.energy_end <= st.energy_start
        assert abs(st.iterations - it_o) <= 2, (step, st.iterations, it_o)
        p0 = (g_before[:, :1] * g_before[:, 4:7]).sum(0, dtype=np.float64)
        p1 = (g_after[:, :1] * g_after[:, 4:7]).sum(0, dtype=np.float64)
        dp = np.abs(g_after[:, :1] * (g_after[:, 4:7] - g_before[:, 4:7])).sum(dtype=np.float64)
        # internal forces only: the exact minimiser conserves momentum (sum_i grad w_ip = 0); the optimiser stops at |step| < 1e-2
        # in an un-weighted norm, which leaves a few per cent of the exchanged impulse
        assert np.abs(p1 - p0).max() <= 0.1 * dp + 1e-7 * np.abs(p0).max(), (p0, p1, dp)
        go = o.grid()
        used = np.flatnonzero(go[:, 0] != 0)
        moved = np.abs(go[used][:, 4:7] - go_before[used][:, 4:7]).max()
        dv = np.abs(g_after[used][:, 4:7] - go[used][:, 4:7]).max()
        assert dv <= 5e-2 * moved + 2e-5 * np.abs(go[used][:, 4:7]).max(), (step, dv, moved)
        o.collisions(dt, ocols, onc); o.fupdate(dt); o.g2p(); o.advect(dt)
        sim.gridBasedCollisions(dt, cols, nc); sim.updateDeformationGradient(dt); sim.updateParticleVelocities(); sim.updateParticlePositions(dt)
    assert total_iters > 2                                                      # the solves did real work
    a, b = sim.download_state35(), o.state()
    assert np.abs(a[:, 5:8] - b[:, 5:8]).max() < 2e-4                           # positions [m]: two steps of dt = 5e-5 at a velocity agreement of ~2 m/s (h = 0.05)
    assert np.abs(a[:, 1:4] - b[:, 1:4]).max() < 5e-2 * 200.0                   # velocities: 5 % of the impact speed


def _crowded_scene(n, seed=3):
    """n particles inside ONE 4^3-cell particle block (several chunks of the tile scatter, ragged 32-slices in the gather) plus a
    few isolated ones, on a 32^3 grid (the scene of test_gpu_parity.test_ragged_block_occupancy_vs_oracle)."""
    rng = np.random.default_rng(seed)
    lo = np.float32(13 * 0.05 + 1e-4)
    pos = (lo + rng.uniform(0, 4 * 0.05 - 2e-4, size=(n, 3))).astype(np.float32)
    pos = np.concatenate([pos, np.array([[0.31, 0.52, 0.77], [1.2, 0.4, 0.9], [0.8, 1.3, 0.2]], np.float32)])
    vel = rng.normal(scale=5.0, size=pos.shape).astype(np.float32)
    sc = mpm_b200.scenes.small_ball(grid=32, radius_cells=2.0, with_ground=False)
    sc.update(pos=pos, vel=vel, mass=np.full(len(pos), 6e-5, np.float32), n=len(pos))
    return sc


@pytest.mark.parametrize("n", [1, 31, 700, 3000])
def test_objective_on_ragged_block_occupancy_vs_oracle(n):
    # energy and gradient with everything in one particle block (1 .. 6 chunks of 512 in the tile scatter) against the oracle,
    # at trial velocities away from the grid's own
    sc = _crowded_scene(n)
    o, _, _ = oracle_from_scene(sc)
    sim, _, _ = sim_from_scene(sc)
    o.rasterize(); sim.rasterizeParticlesToGrid()
    used = o.used_cells()
    rng = np.random.default_rng(11)
    pert = np.zeros((32 ** 3, 3), np.float32)
    pert[used] = rng.standard_normal((used.size, 3)).astype(np.float32)
    dt = 2e-4
    qo, qg = op.default_implicit_params(mu0=50.0, lambda0=30.0), mpm_b200.capi.default_implicit_params(mu0=50.0, lambda0=30.0)
    v = o.grid()[used][:, 4:7] + pert[used]
    e_o = o.energy(v, dt, qo)
    e_g, _ = sim.energy(dt, pert, relative=True, params=qg)
    assert abs(e_g - e_o) <= 2e-4 * abs(e_o), (n, e_g, e_o)          # |F - R| ~ 1e-2 here: the reference's float F - R carries ~1e-5 relative noise, the device's is double
    g_o = o.energy_gradient(v, dt, qo)
    g_g = sim.energy_gradient(dt, pert, relative=True, params=qg)[used]
    assert np.abs(g_g - g_o).max() <= 5e-4 * np.abs(g_o).max(), (n, np.abs(g_g - g_o).max(), np.abs(g_o).max())      # same float-vs-double F - R noise; a lost chunk would be O(1)


def test_time_integration_full_size_ball_vs_oracle():
    # 58 k particles (a 12-cell-radius ball at 200 m/s, 48^3 grid, ~190 particle blocks), three optimiser iterations with the
    # explicit path's material at 10 x the explicit time step, against the oracle running the same three iterations. The
    # reference's stopping rule (|grad| < 1e-2, mathy.hpp:31) would stop at once in these units, so it is tightened on both sides.
    sc = mpm_b200.scenes.small_ball(grid=48, radius_cells=12.0)
    dt = 1e-4
    E, nu = 1.4e5, 0.2
    kw = dict(mu0=E / (2 * (1 + nu)), lambda0=E * nu / ((1 + nu) * (1 - 2 * nu)), hardening=1, max_iters=3, tol_grad=1e-9, tol_step=1e-9)
    o, _, _ = oracle_from_scene(sc)
    sim, _, _ = sim_from_scene(sc)
    st = sim.download_state35()
    st[:, 12] *= 0.99                                      # a uniform 1 % compression along y in FE: something to solve
    o.set_state(st); sim.upload_state35(st)
    o.rasterize(); o.grid_velocities(dt)
    sim.rasterizeParticlesToGrid(); sim.gridVelocitiesUpdate(dt)
    used = o.used_cells()
    vo_star, vg_star = o.grid()[used][:, 4:7].copy(), sim.grid()[used][:, 4:7].copy()
    it_o, _ = o.time_integration(dt, op.default_implicit_params(**kw))
    stt = sim.timeIntegration(dt, mpm_b200.capi.default_implicit_params(**kw))
    d_o, d_g = o.grid()[used][:, 4:7] - vo_star, sim.grid()[used][:, 4:7] - vg_star          # what each solve did to its own v*
    assert stt.iterations == 3 and it_o == 3
    moved = np.abs(d_o).max()
    assert moved > 1e-3 and stt.energy_end < stt.energy_start
    # velocities near 200 m/s carry 1.5e-5 per ulp, so differences of two of them resolve the move to ~3e-5
    assert np.abs(d_g - d_o).max() <= 5e-2 * moved + 6e-5, (np.abs(d_g - d_o).max(), moved)
    m = o.grid()[used][:, :1]
    assert np.abs((m * d_o).sum(0) - (m * d_g).sum(0)).max() <= 5e-2 * np.abs(m * d_o).sum() + 1e-9


def test_implicit_error_paths():
    sc = mpm_b200.scenes.small_ball(grid=32, radius_cells=3.0)
    sim, cols, nc = sim_from_scene(sc)
    sim.substep(float(sc["dt"]), cols, nc, 1)                                  # the re-sorting gather leaves the handle un-binned
    with pytest.raises(mpm_b200.capi.MpmError, match="rasterize"):
        sim.timeIntegration(1e-4)
    sim.rasterizeParticlesToGrid()
    with pytest.raises(mpm_b200.capi.MpmError, match="bad implicit parameters"):
        sim.timeIntegration(1e-4, mpm_b200.capi.default_implicit_params(ls_tau=1.5))
    st = sim.timeIntegration(1e-4, mpm_b200.capi.default_implicit_params(max_iters=0))      # zero iterations: velocities untouched
    assert st.iterations == 0 and st.evaluations >= 1
    simq, _, _ = sim_from_scene(sc, stencil=1)
    with pytest.raises(mpm_b200.capi.MpmError, match="cubic"):
        simq.timeIntegration(1e-4)
