"""Small fused + staged run for compute-sanitizer (memcheck / racecheck / initcheck).
  python tools/sanitize_case.py [p2g_variant g2p_variant]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import mpm_b200
sc = mpm_b200.scenes.small_ball(grid=32, radius_cells=4.0)
pv, gv = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 0)
p = mpm_b200.capi.default_params(p2g_variant=pv, g2p_variant=gv)
sim = mpm_b200.Sim(32, 32, 32, sc["n"], p)
sim.upload(sc["pos"], sc["vel"], sc["mass"])
sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()
cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
sim.staged_substep(1e-5, cols, nc)
sim.substep(1e-5, cols, nc, 3)
# implicit time integration (mpm_implicit.cuh): gradient gather, stress, tile scatter, the optimiser's vector kernels
# (a strained state, so that the optimiser really iterates: history columns, recursion scalars on the device, line search)
s35 = sim.download_state35()
s35[:, 8] *= 1.01; s35[:, 12] *= 0.99; s35[:, 9] += 0.004
sim.upload_state35(s35)
sim.rasterizeParticlesToGrid(); sim.gridVelocitiesUpdate(1e-4)
st = sim.timeIntegration(1e-4, mpm_b200.capi.default_implicit_params(mu0=5.8e4, lambda0=3.9e4, hardening=1, max_iters=12, tol_step=1e-7, tol_grad=1e-7))
assert st.iterations >= 3 and st.energy_end < st.energy_start, (st.iterations, st.energy_start, st.energy_end)
print("implicit", st.iterations, st.evaluations, st.energy_start, st.energy_end)
sim.gridBasedCollisions(1e-4, cols, nc); sim.updateDeformationGradient(1e-4); sim.updateParticleVelocities(); sim.updateParticlePositions(1e-4)
s = sim.download_state35()
print("ok", np.isfinite(s).all(), sim.stats().n_active_nodes)
