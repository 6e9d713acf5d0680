"""Per-phase wall-clock split of the block-tile P2G kernel (clock64 in thread 0 of every CTA, summed). Needs a library built
with -DMPM_P2G_PROFILE:  python tools/p2g_phase_profile.py build [extra nvcc -D flags]  then, on the GPU box,
  MPM_B200_LIB=realtime-deformations_b200/libmpm_b200_prof.so python tools/p2g_phase_profile.py run 512 67108864 p2g:g2p"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "build":
    out = os.path.join(ROOT, "realtime-deformations_b200", sys.argv[2])
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared"] + sys.argv[3:] + ["-o", out, os.path.join(ROOT, "realtime-deformations_b200", "csrc", "mpm_api.cu")]
    subprocess.check_call(cmd); print(out); sys.exit(0)
import mpm_b200
grid, n = int(sys.argv[2]), int(sys.argv[3])
pv, gv = (int(x) for x in sys.argv[4].split(":"))
sc = mpm_b200.scenes.snow_slab(grid=grid, n=n)
p = mpm_b200.capi.default_params(p2g_variant=pv, g2p_variant=gv)
p.gravity[:] = [float(x) for x in sc["gravity"]]
sim = mpm_b200.Sim(grid, grid, grid, sc["n"], p)
sim.upload(sc["pos"], sc["vel"], sc["mass"]); sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()
cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
sim.substep(1e-5, cols, nc, 3)
t = (C.c_int64 * 8)()
sim.L.mpm_debug_p2g_profile(sim.h, t, 1)
steps = 5
sim.substep(1e-5, cols, nc, steps)
sim.L.mpm_debug_p2g_profile(sim.h, t, 0)
tot = sum(t) or 1
names = ["top wait", "derive", "sort", "accumulate", "barrier after acc", "z-fold+stores", "xy-fold+reds", "F-update"]
ms = list(sim.stats().last_ms)
print(f"variants {pv}:{gv}  p2g kernel {ms[2]:.3f} ms  (lib {os.environ.get('MPM_B200_LIB', 'default')})")
for nm, v in zip(names, t):
    print(f"  {nm:20s} {100.0 * v / tot:5.1f} %   {v / tot * ms[2]:.3f} ms")
