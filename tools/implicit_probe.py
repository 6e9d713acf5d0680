"""Timing probe of the implicit time integration (not the contract bench):
  python tools/implicit_probe.py GRID N [dt] [explicit substeps first]
One semi-implicit substep on the config-2 style snowball (rasterize | gravity | mpm_time_integration | collisions | F-update |
G2P | advect), with the material of the explicit path; prints the optimiser's statistics and the device time per objective
evaluation (one pass over the particles with gradient scatter)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import mpm_b200

grid, n = int(sys.argv[1]), int(sys.argv[2])
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
sc = mpm_b200.scenes.snowball_drop(grid=grid, n=n)
import os
form = int(os.environ.get("MPM_PROBE_BASELINE", "0"))      # 1 = thread-per-particle baseline kernels instead of the block-tile form
p = mpm_b200.capi.default_params(p2g_variant=form, g2p_variant=form)
sim = mpm_b200.Sim(grid, grid, grid, sc["n"], p)
sim.upload(sc["pos"], sc["vel"], sc["mass"])
sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()
cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
pre = int(sys.argv[4]) if len(sys.argv) > 4 else 400
sim.substep(1e-5, cols, nc, pre)            # explicit substeps until the ball is well into the impact
E, nu = 1.4e5, 0.2
q = mpm_b200.capi.default_implicit_params(mu0=E / (2 * (1 + nu)), lambda0=E * nu / ((1 + nu) * (1 - 2 * nu)), hardening=1)
for step in range(3):
    sim.rasterizeParticlesToGrid(); sim.gridVelocitiesUpdate(dt); sim.synchronize()
    t0 = time.perf_counter(); st = sim.timeIntegration(dt, q); sim.synchronize(); t = time.perf_counter() - t0
    sim.gridBasedCollisions(dt, cols, nc); sim.updateDeformationGradient(dt); sim.updateParticleVelocities(); sim.updateParticlePositions(dt)
    print(f"step {step}: form={'baseline' if form else 'tile'} n={sc['n']} grid={grid} dt={dt} iterations={st.iterations} evaluations={st.evaluations} E {st.energy_start:.6g} -> {st.energy_end:.6g} |grad| {st.grad_norm_end:.3g} "
          f"solve {t*1e3:.2f} ms = {t*1e3/max(st.evaluations,1):.3f} ms per evaluation, {sc['n']*st.evaluations/t/1e9:.3f} G particle-evaluations/s", flush=True)
# evaluation alone (energy + gradient), timed over 10 calls
sim.rasterizeParticlesToGrid(); sim.synchronize()
z = np.zeros((grid ** 3, 3), np.float32)
t0 = time.perf_counter()
for _ in range(5):
    sim.energy(dt, None)
sim.synchronize()
print(f"mpm_energy (value only, incl. host read-back): {(time.perf_counter()-t0)/5*1e3:.3f} ms")
