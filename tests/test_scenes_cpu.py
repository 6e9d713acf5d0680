"""CPU tests of the synthetic scene generators (host-side numpy; SURVEY 8(d) inputs for BASELINE configs 2-5)."""
import numpy as np

import mpm_b200

S = mpm_b200.scenes


def _inside(sc):
    h, dims = float(sc["h"]), np.asarray(sc["dims"])
    cell = (sc["pos"] / np.float32(h)).astype(np.int32)
    return (cell >= 3).all() and (cell < dims - 3).all()          # inside the reference's clamp box (cpp:381-388)


def test_every_cell_of_a_filled_region_holds_eight_particles():
    """The fill rule of initializeParticles (cpp:29-54): 8 jittered sites per cell that stay inside their cell -- the
    layout the benchmark scene has and P2G's rotated record walk is built for."""
    p = S.box_region((4, 5, 6), (9, 8, 12), (32, 32, 32), 0.05)
    cell = (p / np.float32(0.05)).astype(np.int32)
    key = (cell[:, 0] * 64 + cell[:, 1]) * 64 + cell[:, 2]
    _, counts = np.unique(key, return_counts=True)
    assert len(p) == 5 * 3 * 6 * 8 and (counts == 8).all()
    assert cell.min(0).tolist() == [4, 5, 6] and cell.max(0).tolist() == [8, 7, 11]


def test_generators_are_reproducible_and_cell_keyed():
    a, b = S.ball((1.0, 1.0, 1.0), 0.3, (40, 40, 40), 0.05), S.ball((1.0, 1.0, 1.0), 0.3, (40, 40, 40), 0.05)
    assert np.array_equal(a, b)
    part = S.ball((1.0, 1.0, 1.0), 0.3, (40, 40, 40), 0.05, i_range=(18, 22))      # a slab of the same ball: same particles
    ci = (a[:, 0] / np.float32(0.05)).astype(np.int32)
    assert np.array_equal(part, a[(ci >= 18) & (ci < 22)])
    assert not np.array_equal(S.ball((1.0, 1.0, 1.0), 0.3, (40, 40, 40), 0.05, seed=S.SEED + 1)[:100], a[:100])


def test_config_scenes_have_the_requested_size_and_fit_the_grid():
    drop = S.snowball_drop(grid=32, n=4096)
    two = S.snowball_collision(grid=64, n=1 << 13)
    slab = S.snow_slab(grid=32, n=8192)
    stiff = S.stiff_snowball(grid=32, n=4096)
    for sc, n in ((drop, 4096), (two, 1 << 13), (slab, 8192), (stiff, 4096)):
        assert sc["n"] == n == len(sc["pos"]) == len(sc["vel"]) == len(sc["mass"]) and _inside(sc)
        assert (sc["mass"] == np.float32(0.00006)).all()                          # cpp:53
    # config 3: two balls approaching each other along i, no ground
    half = two["n"] // 2
    assert (two["vel"][:half, 0] == 100).all() and (two["vel"][half:, 0] == -100).all() and len(two["w2l"]) == 0
    assert two["pos"][:half, 0].max() < two["pos"][half:, 0].min()
    # config 4: small dt, ball a quarter of a cell above the ground box
    top = 0.125 * 32 * 0.05 + 0.025
    assert stiff["dt"] == np.float32(2.5e-6) and 0.0 < stiff["pos"][:, 1].min() - top < 1.0 * 0.05
    # config 5: tilted gravity, at rest, 8 particles in (nearly) every occupied cell
    assert abs(float(np.linalg.norm(slab["gravity"])) - 9.8) < 1e-5 and slab["gravity"][0] > 0 and (slab["vel"] == 0).all()
    cell = (slab["pos"] / np.float32(0.05)).astype(np.int32)
    _, counts = np.unique((cell[:, 0] * 64 + cell[:, 1]) * 64 + cell[:, 2], return_counts=True)
    assert (counts == 8).mean() > 0.99


def test_slab_layer_counts_match_the_generated_scene():
    grid, n = 64, 100000
    sc = S.snow_slab(grid=grid, n=n)
    layer = ((sc["pos"][:, 0] / np.float32(0.05)).astype(np.int32) - 1) >> 2
    assert np.array_equal(np.bincount(layer, minlength=(grid + 3) // 4), S.snow_slab_layer_counts(grid, n))
