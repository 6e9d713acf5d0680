"""TEST INFRASTRUCTURE — regenerates tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/ref_mpm (reference material_point_method.cpp + mathy.cpp compiled where they lie by
oracle/Makefile, driven by oracle/ref_driver.cpp) and packs its raw float32 dumps. Needs /root/reference,
so it only runs in the build container; the .npz files are committed and are what travels.

    make -C oracle ref && python oracle/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "ref_mpm")
OUT = os.path.join(HERE, "..", "tests", "golden")
SEED = 20260117


def run(args):
    r = subprocess.run([REF] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.exit("ref_mpm failed: " + r.stdout)
    return r.stdout


def ld(path, w):
    return np.fromfile(path, dtype=np.float32).reshape(-1, w)


def sparse_grid(g, cols):
    used = np.flatnonzero(g[:, 0] != 0).astype(np.int32)
    return used, g[used][:, cols]


def c1_default(tmp):
    """Config 1: the reference's default scene (main.cpp:48-52,119-156), trajectories + stage-level dumps."""
    traj = [1, 20, 100, 200, 400]
    stages = [1, 120]
    run(["--steps", max(traj), "--dump-dir", tmp, "--dump-steps", ",".join(map(str, traj)),
         "--stage-steps", ",".join(map(str, stages)), "--quiet"])
    out = {"I": 20, "J": 20, "K": 20, "h": np.float32(0.05), "dt": np.float32(1e-5),
           "colliders": np.fromfile(tmp + "/colliders.f32", dtype=np.float32).reshape(-1, 29),
           "state0": ld(tmp + "/particles_step0000.f32", 35), "traj_steps": np.array(traj),
           "stage_steps": np.array(stages)}
    for n in traj:
        out[f"state{n}"] = ld(f"{tmp}/particles_step{n:04d}.f32", 35)
    for s in stages:
        pre = ld(f"{tmp}/stage{s:04d}_pre.f32", 35)
        out[f"st{s}_pre"] = pre
        g = ld(f"{tmp}/stage{s:04d}_p2g.f32", 7)
        used, vals = sparse_grid(g, [0, 4, 5, 6])
        out[f"st{s}_used"] = used
        out[f"st{s}_p2g"] = vals                                      # mass, velocity
        out[f"st{s}_forces"] = ld(f"{tmp}/stage{s:04d}_forces.f32", 7)[used][:, 1:4]
        out[f"st{s}_gridvel"] = ld(f"{tmp}/stage{s:04d}_gridvel.f32", 7)[used][:, 4:7]
        out[f"st{s}_collide"] = ld(f"{tmp}/stage{s:04d}_collide.f32", 7)[used][:, 4:7]
        out[f"st{s}_fupdate"] = ld(f"{tmp}/stage{s:04d}_fupdate.f32", 35)[:, 8:26]   # FE, FP
        out[f"st{s}_g2p"] = ld(f"{tmp}/stage{s:04d}_g2p.f32", 35)[:, [1, 2, 3] + list(range(26, 35))]  # vel, B
        out[f"st{s}_advect"] = ld(f"{tmp}/stage{s:04d}_advect.f32", 35)[:, 5:8]
    np.savez_compressed(os.path.join(OUT, "c1_default.npz"), **out)


def c1b_fine(tmp):
    """h = 0.025, 40^3, 17 100 particles: the largest scene the real class can hold. Snapshots in free
    fall (20 substeps) and after first contact with the ground box (120 substeps)."""
    n, steps, late = 17100, 20, 120
    run(["--h", 0.025, "--grid", 40, 40, 40, "--n", n, "--steps", late, "--dump-dir", tmp,
         "--dump-steps", f"{steps},{late}", "--quiet"])
    s0 = ld(tmp + "/particles_step0000.f32", 35)
    s1 = ld(f"{tmp}/particles_step{steps:04d}.f32", 35)
    s2 = ld(f"{tmp}/particles_step{late:04d}.f32", 35)
    np.savez_compressed(os.path.join(OUT, "c1b_h0025.npz"), I=40, J=40, K=40, h=np.float32(0.025), dt=np.float32(1e-5),
                        steps=steps, colliders=np.fromfile(tmp + "/colliders.f32", dtype=np.float32).reshape(-1, 29),
                        pos0=s0[:, 5:8], vel0=s0[0, 1:4], mass0=s0[0, 0], volume0=s0[:, 4],
                        pos=s1[:, 5:8], vel=s1[:, 1:4], FE=s1[:, 8:17], FP=s1[:, 17:26],
                        steps_late=late, pos_late=s2[:, 5:8], vel_late=s2[:, 1:4], FE_late=s2[:, 8:17], FP_late=s2[:, 17:26])


def rand_rot(rng, n):
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)


def kats(tmp):
    rng = np.random.default_rng(SEED)
    out = {}
    # --- weightNx (material_point_method.hpp:20-31)
    x = np.concatenate([np.linspace(-2.5, 2.5, 2001), rng.uniform(-2.2, 2.2, 4096),
                        [0.0, -0.0, 1.0, -1.0, 2.0, -2.0, np.nextafter(np.float32(1), np.float32(0)),
                         np.nextafter(np.float32(1), np.float32(2)), np.nextafter(np.float32(2), np.float32(0)),
                         np.nextafter(np.float32(2), np.float32(3)), 1e-30, 3.0, 100.0]]).astype(np.float32)
    x.tofile(tmp + "/w_in.f32")
    run(["--kat-weights", tmp + "/w_in.f32", tmp + "/w_out.f32"])
    out["weights_x"] = x
    out["weights_w"] = np.fromfile(tmp + "/w_out.f32", dtype=np.float32)
    # --- polarDecomposition (utils.h:55-75), FE-like inputs: rotation x stretch in [0.9, 1.1] + shear noise
    n = 2048
    R = rand_rot(rng, n)
    sig = rng.uniform(0.9, 1.1, size=(n, 3)); sig[: n // 2] = rng.uniform(0.975, 1.005, size=(n // 2, 3))
    Q = rand_rot(rng, n)
    S = Q @ (sig[:, :, None] * np.transpose(Q, (0, 2, 1)))
    F = (R @ S).astype(np.float32)
    F[0] = np.eye(3)
    Fg = np.transpose(F, (0, 2, 1)).reshape(n, 9).copy()        # glm column-major
    Fg.tofile(tmp + "/p_in.f32")
    run(["--kat-polar", tmp + "/p_in.f32", tmp + "/p_out.f32"])
    po = np.fromfile(tmp + "/p_out.f32", dtype=np.float32).reshape(n, 18)
    out["polar_F"] = Fg
    out["polar_R"] = po[:, :9]
    # --- bodyCollision (material_point_method.cpp:264-296) on every node of the default 20^3 grid
    ii, jj, kk = np.meshgrid(np.arange(20), np.arange(20), np.arange(20), indexing="ij")
    pos = (np.stack([ii, jj, kk], -1).reshape(-1, 3).astype(np.float32) * np.float32(0.05)).astype(np.float32)
    extra = rng.uniform(0.0, 1.0, size=(4000, 3)).astype(np.float32)      # off-node points as well
    pos = np.concatenate([pos, extra])
    vel = rng.normal(scale=50.0, size=pos.shape).astype(np.float32)
    vel[::7] = np.array([0.0, -200.0, 0.0], np.float32)
    vel[5::11, 1] *= 20.0                                                 # some with -mu*vn > 3 (sticking branch)
    np.concatenate([pos, vel], 1).astype(np.float32).tofile(tmp + "/c_in.f32")
    run(["--kat-collide", tmp + "/c_in.f32", tmp + "/c_out.f32", "--dump-dir", tmp])
    out["collide_pos"] = pos
    out["collide_vel"] = vel
    out["collide_out"] = np.fromfile(tmp + "/c_out.f32", dtype=np.float32).reshape(-1, 3)
    out["colliders"] = np.fromfile(tmp + "/colliders.f32", dtype=np.float32).reshape(-1, 29)
    # same points against MOVING colliders (MeshCollider::velocity != 0, the key_callback case of main.cpp:37-41,174-190)
    run(["--kat-collide", tmp + "/c_in.f32", tmp + "/cm_out.f32", "--dump-dir", tmp, "--collider-vel", 3.0, -1.5, 0.75])
    out["collide_moving_out"] = np.fromfile(tmp + "/cm_out.f32", dtype=np.float32).reshape(-1, 3)
    out["colliders_moving"] = np.fromfile(tmp + "/colliders.f32", dtype=np.float32).reshape(-1, 29)
    # --- updateDeformationGradient (material_point_method.cpp:306-330) on synthetic particle states
    n = 4096
    st = np.zeros((n, 35), np.float32)
    st[:, 0] = 6e-5; st[:, 4] = 3e-5; st[:, 5:8] = 0.5
    R = rand_rot(rng, n); Q = rand_rot(rng, n)
    sig = rng.uniform(0.975, 1.005, size=(n, 3))
    FE = R @ (Q @ (sig[:, :, None] * np.transpose(Q, (0, 2, 1))))
    Q2 = rand_rot(rng, n); R2 = rand_rot(rng, n)
    sp = np.exp(rng.uniform(-0.7, 0.7, size=(n, 3)))
    FP = R2 @ (Q2 @ (sp[:, :, None] * np.transpose(Q2, (0, 2, 1))))
    Bm = rng.normal(size=(n, 3, 3)) * np.exp(rng.uniform(np.log(1e-4), np.log(60.0), size=(n, 1, 1)))
    # special cases: identity / zero B, diagonal F with exact zeros off-diagonal, repeated singular values,
    # reflections (negative diagonal), permutation-like matrices
    FE[0] = np.eye(3); FP[0] = np.eye(3); Bm[0] = 0
    FE[1] = np.diag([1.0, 1.0, 1.0]); FP[1] = np.eye(3); Bm[1] = np.diag([10.0, -20.0, 5.0])
    FE[2] = np.diag([0.99, 0.99, 0.99]); FP[2] = np.eye(3); Bm[2] = 0
    FE[3] = np.diag([-1.0, 1.0, 1.0]); FP[3] = np.eye(3); Bm[3] = 0
    FE[4] = np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1.0]]); FP[4] = np.eye(3); Bm[4] = 0
    FE[5] = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0.0]]); FP[5] = np.eye(3); Bm[5] = np.full((3, 3), 1.0)
    FE[6] = np.diag([1.002, 0.98, 0.98]); FP[6] = np.diag([1.0, 1.1, 0.9]); Bm[6] = 0
    FE[7] = np.eye(3); FP[7] = np.eye(3); Bm[7] = np.array([[0, 30.0, 0], [-30.0, 0, 0], [0, 0, 0]])
    st[:, 8:17] = np.transpose(FE, (0, 2, 1)).reshape(n, 9)
    st[:, 17:26] = np.transpose(FP, (0, 2, 1)).reshape(n, 9)
    st[:, 26:35] = np.transpose(Bm, (0, 2, 1)).reshape(n, 9)
    st.tofile(tmp + "/f_in.f32")
    run(["--n", n, "--load-full", tmp + "/f_in.f32", "--kat-fupdate", tmp + "/f_out.f32"])
    fo = ld(tmp + "/f_out.f32", 35)
    out["fupdate_in"] = st[:, 8:35]          # FE, FP, B
    out["fupdate_out"] = fo[:, 8:26]         # FE, FP
    np.savez_compressed(os.path.join(OUT, "kat_functions.npz"), **out)


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle ref  (needs /root/reference)")
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as t:
        c1_default(t)
    with tempfile.TemporaryDirectory() as t:
        c1b_fine(t)
    with tempfile.TemporaryDirectory() as t:
        kats(t)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
