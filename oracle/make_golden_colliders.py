"""TEST INFRASTRUCTURE — tests/golden/kat_colliders.npz: MeshCollider poses straight from the UNMODIFIED reference.

oracle/_ref/ref_mpm builds boxes exactly like main.cpp:119-151 (applyMatrix4(translate * rotate_z * scale), which runs
glm::decompose) from random rows, gives them a velocity, calls MeshCollider::move (hpp:90-92) once per substep, and dumps
for every box what its sdf lambda uses (hpp:80-83: scale, rotation quaternion, translation) together with
inverse(translate(translation) * toMat4(rotation)) -- before the first and after the last substep. These vectors pin
the scene front-end of the C ABI (mpm_box_collider_from_transform, mpm_box_transform_move).

    make -C oracle ref && python oracle/make_golden_colliders.py
"""
import os
import tempfile

import numpy as np

from make_golden import OUT, SEED, run

N_BOXES, STEPS, VEL = 24, 25, (3.0, -1.5, 0.75)


def main():
    rng = np.random.default_rng(SEED + 7)
    rows = np.concatenate([rng.uniform(-0.5, 1.5, (N_BOXES, 3)), rng.uniform(-180.0, 180.0, (N_BOXES, 1)),
                           rng.uniform(0.05, 0.6, (N_BOXES, 3))], 1).astype(np.float32)
    rows[0, 3] = 0.0            # axis-aligned and the reference's own 45 degree box among them
    rows[1, 3] = 45.0
    with tempfile.TemporaryDirectory() as tmp:
        rows.tofile(tmp + "/boxes.f32")
        run(["--colliders", tmp + "/boxes.f32", "--collider-vel", *VEL, "--steps", STEPS, "--dump-dir", tmp, "--quiet"])
        first = np.fromfile(tmp + "/colliders.f32", dtype=np.float32).reshape(-1, 29)
        last = np.fromfile(tmp + "/colliders_final.f32", dtype=np.float32).reshape(-1, 29)
    assert first.shape == last.shape == (N_BOXES, 29)
    np.savez_compressed(os.path.join(OUT, "kat_colliders.npz"), rows=rows, first=first, last=last, steps=np.int32(STEPS),
                        dt=np.float32(1e-5), velocity=np.asarray(VEL, np.float32))
    print("kat_colliders.npz:", first.shape, "max |translation drift|", np.abs(last[:, 7:10] - first[:, 7:10]).max())


if __name__ == "__main__":
    main()
