"""Helpers shared by the GPU parity tests: run the same scene through the CUDA library and the CPU oracle."""
import numpy as np

import mpm_b200
import oracle_py as op

import os

VARIANTS = [(0, 0), (1, 1)]     # (p2g_variant, g2p_variant): tile kernels (F-update inside P2G in the fused substep), baseline kernels
if os.environ.get("MPM_TEST_EXPERIMENTAL") == "1":
    VARIANTS += [(2, 0), (0, 1), (1, 0)]      # F-update as its own kernel; mixed tile / baseline pairs
TILE_VARIANTS = [v for v in VARIANTS if v != (1, 1)]     # the tile kernels only (edge cases of block occupancy)


def oracle_from_scene(sc, fma=False, threads=1, **prm):
    I, J, K = sc["dims"]
    p = op.default_params(h=float(sc["h"]), **prm)
    if "gravity" in sc and "gravity" not in prm:
        p.gravity[:] = [float(x) for x in sc["gravity"]]
    o = op.Oracle(I, J, K, sc["n"], p, threads=threads, fma=fma)
    s = op.initial_state(sc["pos"], sc["vel"], sc["mass"])
    o.set_state(s)
    o.rasterize(); o.volumes()
    cols, nc = op.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
    return o, cols, nc


def sim_from_scene(sc, variants=(0, 0), **prm):
    I, J, K = sc["dims"]
    p = mpm_b200.capi.default_params(h=float(sc["h"]), p2g_variant=variants[0], g2p_variant=variants[1], **prm)
    if "gravity" in sc and "gravity" not in prm:
        p.gravity[:] = [float(x) for x in sc["gravity"]]
    sim = mpm_b200.Sim(I, J, K, sc["n"], p)
    sim.upload(sc["pos"], sc["vel"], sc["mass"])
    sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()     # main.cpp:53-54
    cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
    return sim, cols, nc


def sim_from_state35(state, dims, variants=(0, 0), **prm):
    p = mpm_b200.capi.default_params(p2g_variant=variants[0], g2p_variant=variants[1], **prm)
    sim = mpm_b200.Sim(dims[0], dims[1], dims[2], state.shape[0], p)
    sim.upload_state35(state)
    return sim


def gpu_colliders_from_ref_dump(raw):
    raw = np.asarray(raw, np.float32).reshape(-1, 29)
    return mpm_b200.capi.make_colliders(raw[:, 13:29], raw[:, 0:3], raw[:, 10:13])
