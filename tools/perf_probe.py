"""Quick per-phase timing probe (not the contract bench):
  python tools/perf_probe.py GRID N STEPS [slab|ball] [p2g:g2p,p2g:g2p,...]
e.g. A/B of the experimental kernels:  python tools/perf_probe.py 256 8388608 20 slab 0:0,0:2,0:3,0:4,2:0,2:4
A third field switches P2G's record walk: 0:0:0 = aligned (the kernel measured in round 1), 0:0:1 = rotated (default)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import mpm_b200

_scene_cache = {}


def run(grid, n, steps, pv, gv, scene="slab", rotate=None):
    import os
    if rotate is None:
        os.environ.pop("MPM_B200_P2G_ROTATE", None)
    else:
        os.environ["MPM_B200_P2G_ROTATE"] = str(rotate)        # read by mpm_create
    t0 = time.time()
    key = (grid, n, scene)
    if key not in _scene_cache:
        _scene_cache.clear()
        _scene_cache[key] = mpm_b200.scenes.snow_slab(grid=grid, n=n) if scene == "slab" else mpm_b200.scenes.snowball_drop(grid=grid, n=n)
    sc = _scene_cache[key]
    tg = time.time() - t0
    fx = int(os.environ.get("MPM_PROBE_FUPDATE_EXACT", "0"))
    stencil = int(os.environ.get("MPM_PROBE_STENCIL", "0"))          # 1 = quadratic B-spline (not the reference's stencil)
    p = mpm_b200.capi.default_params(p2g_variant=pv, g2p_variant=gv, fupdate_exact=fx, stencil=stencil)
    if "gravity" in sc: p.gravity[:] = [float(x) for x in sc["gravity"]]
    sim = mpm_b200.Sim(grid, grid, grid, sc["n"], p)
    if os.environ.get("MPM_PROBE_SHUFFLE") == "1":          # worst case for everything that relies on cell-coherent particle order
        perm = np.random.default_rng(1).permutation(sc["n"])
        sc = dict(sc, pos=sc["pos"][perm], vel=sc["vel"][perm], mass=sc["mass"][perm])
    t0 = time.time(); sim.upload(sc["pos"], sc["vel"], sc["mass"]); tu = time.time() - t0
    sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()
    cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
    sim.substep(1e-5, cols, nc, 3); sim.synchronize()
    t0 = time.time(); sim.substep(1e-5, cols, nc, steps); sim.synchronize(); dt = (time.time() - t0) / steps
    st = sim.stats()
    ms = list(st.last_ms)
    print(f"{scene} grid={grid} n={sc['n']} fexact={fx} stencil={stencil} variants=({pv},{gv}) rotate={'default' if rotate is None else rotate} gen={tg:.1f}s upload={tu:.1f}s  {dt*1e3:.3f} ms/substep  "
          f"{sc['n']/dt/1e9:.3f} G upd/s  bin={ms[0]:.3f} clear={ms[1]:.3f} p2g={ms[2]:.3f} grid={ms[3]:.3f} g2p={ms[4]:.3f} (fupdate={ms[7]:.3f}) total={ms[6]:.3f} "
          f"active_nodes={st.n_active_nodes} pblocks={st.n_particle_blocks} gblocks={st.n_grid_blocks}", flush=True)
    sim.close()

if __name__ == "__main__":
    grid, n, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    scene = sys.argv[4] if len(sys.argv) > 4 else "slab"
    combos = ((0, 0), (1, 1), (0, 1), (1, 0))
    if len(sys.argv) > 5:
        combos = tuple(tuple(int(x) for x in c.split(":")) for c in sys.argv[5].split(","))
    for c in combos:              # "p2g:g2p" or "p2g:g2p:rotate" (rotate 0 = the aligned P2G record walk measured in round 1)
        run(grid, n, steps, c[0], c[1], scene, c[2] if len(c) > 2 else None)
