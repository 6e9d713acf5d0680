"""Runs a few fused substeps on a slab scene; used under ncu (python tools/profile_step.py grid n steps)."""
import sys
sys.path.insert(0, ".")
import mpm_b200
grid, n, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sc = mpm_b200.scenes.snow_slab(grid=grid, n=n)
p = mpm_b200.capi.default_params()
p.gravity[:] = [float(x) for x in sc["gravity"]]
sim = mpm_b200.Sim(grid, grid, grid, sc["n"], p)
sim.upload(sc["pos"], sc["vel"], sc["mass"])
sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()
cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
sim.substep(1e-5, cols, nc, steps); sim.synchronize()
print("done", sim.stats().n_active_nodes)
