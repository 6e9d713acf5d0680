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
s = sim.download_state35()
print("ok", np.isfinite(s).all(), sim.stats().n_active_nodes)
