"""Synthetic snow scenes for the benchmark configurations.

Generalises the reference's fill rule (initializeParticles, material_point_method.cpp:29-54): every cell of a
region gets 8 sub-cell sites (cell + {1/4,3/4}^3 + jitter inside a ball of radius 0.25 cell) * h, kept if inside
the body. The jitter comes from a counter-based hash keyed by (seed, global cell id, site, stream), so any
sub-range of cells — e.g. one GPU's slab — can be generated independently and reproducibly. m_p = 6e-5 (cpp:53).
All arrays are created on the HOST (numpy); nothing here touches the GPU.
"""
import numpy as np

SEED = 20260117
PARTICLE_MASS = np.float32(0.00006)
_SITES = np.array([[1, 1, 1], [1, 1, 3], [1, 3, 1], [1, 3, 3], [3, 1, 1], [3, 1, 3], [3, 3, 1], [3, 3, 3]], np.float32) * 0.25


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform(keys, stream, seed):
    with np.errstate(over="ignore"):
        z = _splitmix64(keys * np.uint64(4) + np.uint64(stream) + np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D))
        z = _splitmix64(z)
    return ((z >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))).astype(np.float32)


def fill_cells(i0, i1, j0, j1, k0, k1, dims, h, seed=SEED):
    """Candidate positions ((i1-i0)*(j1-j0)*(k1-k0)*8, 3) float32 in i-major cell order, site fastest."""
    I, J, K = dims
    ii, jj, kk = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), np.arange(k0, k1), indexing="ij")
    cell = np.stack([ii, jj, kk], -1).reshape(-1, 1, 3).astype(np.float32)
    cid = ((ii.astype(np.uint64) * np.uint64(J) + jj.astype(np.uint64)) * np.uint64(K) + kk.astype(np.uint64)).reshape(-1, 1)
    keys = cid * np.uint64(8) + np.arange(8, dtype=np.uint64).reshape(1, 8)
    phi = _uniform(keys, 0, seed) * np.float32(2.0 * 3.1415)
    cost = _uniform(keys, 1, seed) * np.float32(2.0) - np.float32(1.0)
    u = _uniform(keys, 2, seed)
    r = np.float32(0.25) * np.cbrt(u)
    sint = np.sqrt(np.maximum(np.float32(0), np.float32(1) - cost * cost))
    jit = np.stack([r * sint * np.cos(phi), r * sint * np.sin(phi), r * cost], -1).astype(np.float32)
    pos = (cell + _SITES.reshape(1, 8, 3) + jit) * np.float32(h)
    return pos.reshape(-1, 3).astype(np.float32)


def ball(center, radius, dims, h, seed=SEED, n_max=None, i_range=None):
    """Particles of a ball; candidates are produced plane by plane to bound memory. i_range restricts the cells."""
    c = np.asarray(center, np.float32)
    lo = np.maximum(np.floor((c - radius) / h).astype(int) - 1, 0)
    hi = np.minimum(np.floor((c + radius) / h).astype(int) + 2, np.asarray(dims))
    if i_range is not None:
        lo[0], hi[0] = max(lo[0], i_range[0]), min(hi[0], i_range[1])
    out, total = [], 0
    for i in range(lo[0], hi[0]):
        p = fill_cells(i, i + 1, lo[1], hi[1], lo[2], hi[2], dims, h, seed)
        d = p - c
        p = p[np.sqrt((d * d).sum(1)) <= np.float32(radius)]
        out.append(p); total += len(p)
        if n_max is not None and total >= n_max:
            break
    pos = np.concatenate(out) if out else np.zeros((0, 3), np.float32)
    return pos[:n_max] if n_max is not None else pos


def box_region(cell_lo, cell_hi, dims, h, seed=SEED, n_max=None):
    out, total = [], 0
    for i in range(cell_lo[0], cell_hi[0]):
        p = fill_cells(i, i + 1, cell_lo[1], cell_hi[1], cell_lo[2], cell_hi[2], dims, h, seed)
        out.append(p); total += len(p)
        if n_max is not None and total >= n_max:
            break
    pos = np.concatenate(out) if out else np.zeros((0, 3), np.float32)
    return pos[:n_max] if n_max is not None else pos


def ground_collider(top_y, dims, h):
    """Axis-aligned ground box with its top face at top_y (off-node by h/2 in the presets, so sdf != 0 on nodes).
    world_to_local = inverse(translate(t) * identity rotation) as a glm column-major mat4."""
    I, J, K = dims
    half = np.array([I * h, 2.0, K * h], np.float32)
    t = np.array([I * h / 2, top_y - 2.0, K * h / 2], np.float32)
    w2l = np.eye(4, dtype=np.float32)
    w2l[3, 0:3] = -t          # row-major view of a column-major mat4: translation lives in elements 12..14
    return w2l.reshape(16), half, np.zeros(3, np.float32)


def _scene(pos, vel, dims, h, dt, colliders, **extra):
    n = pos.shape[0]
    w2l = np.stack([c[0] for c in colliders]) if colliders else np.zeros((0, 16), np.float32)
    half = np.stack([c[1] for c in colliders]) if colliders else np.zeros((0, 3), np.float32)
    cvel = np.stack([c[2] for c in colliders]) if colliders else np.zeros((0, 3), np.float32)
    d = dict(pos=pos, vel=np.broadcast_to(np.asarray(vel, np.float32), (n, 3)).copy() if np.ndim(vel) == 1 else vel,
             mass=np.full(n, PARTICLE_MASS, np.float32), dims=tuple(dims), h=np.float32(h), dt=np.float32(dt),
             w2l=w2l, half=half, cvel=cvel, n=n)
    d.update(extra)
    return d


def snowball_drop(grid=128, n=1 << 20, h=0.05, dt=1e-5, seed=SEED):
    """Config 2: one snowball dropped on a ground plane. Ball radius = smallest giving >= n candidates."""
    dims = (grid, grid, grid)
    L = grid * h
    top = 0.125 * L + h / 2
    r_cells = (n / 8.0 * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    center = (0.5 * L, top + (r_cells + 4) * h, 0.5 * L)
    radius = r_cells * h
    while True:
        pos = ball(center, radius, dims, h, seed)
        if len(pos) >= n:
            break
        radius *= 1.01
    return _scene(pos[:n], (0.0, -200.0, 0.0), dims, h, dt, [ground_collider(top, dims, h)], name=f"snowball_drop_{grid}")


def stiff_snowball(grid=256, n=1 << 22, h=0.05, dt=2.5e-6, seed=SEED, gap_cells=0.25):
    """Config 4: one snowball at small dt for the stiff-snow parameter sweep (hardening xi, theta_c, theta_s). Same ground
    box as config 2; the ball starts `gap_cells` above it, so that at 200 m/s the first contact comes after
    gap_cells * h / (200 dt) substeps (25 by default) and the plasticity clamps are active within a short run."""
    dims = (grid, grid, grid)
    L = grid * h
    top = 0.125 * L + h / 2
    radius = (n / 8.0 * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0) * h
    while True:
        center = (0.5 * L, top + radius + gap_cells * h, 0.5 * L)
        pos = ball(center, radius, dims, h, seed)
        if len(pos) >= n:
            break
        radius *= 1.01
    return _scene(pos[:n], (0.0, -200.0, 0.0), dims, h, dt, [ground_collider(top, dims, h)], name=f"stiff_snowball_{grid}")


def snowball_collision(grid=256, n=1 << 23, h=0.05, dt=1e-5, seed=SEED, gap_cells=3.6):
    """Config 3: two snowballs colliding head-on along i, no ground. gap_cells = initial surface gap (3.6 cells: first contact
    after ~90 substeps at 2 x 100 m/s; tests that must reach the contact inside a short window pass a smaller gap)."""
    dims = (grid, grid, grid)
    L = grid * h
    half_n = n // 2
    r_cells = (half_n / 8.0 * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    radius = r_cells * h * 1.004
    gap = gap_cells * h
    c1 = (0.5 * L - radius - gap / 2, 0.5 * L, 0.5 * L)
    c2 = (0.5 * L + radius + gap / 2, 0.5 * L, 0.5 * L)
    while True:
        p1, p2 = ball(c1, radius, dims, h, seed), ball(c2, radius, dims, h, seed + 1)
        if len(p1) >= half_n and len(p2) >= half_n:
            break
        radius *= 1.01
    pos = np.concatenate([p1[:half_n], p2[:half_n]])
    vel = np.zeros_like(pos)
    vel[:half_n, 0] = 100.0
    vel[half_n:, 0] = -100.0
    return _scene(pos, vel, dims, h, dt, [], name=f"snowball_collision_{grid}")


def snow_slab(grid=512, n=1 << 26, h=0.05, dt=1e-5, seed=SEED, i_range=None, tilt_deg=30.0):
    """Config 5: slab of snow (8 ppc) resting on a ground box, gravity tilted towards +i so the flow crosses
    slab-decomposition boundaries. Footprint and thickness scale with the grid so that ~n particles fit."""
    dims = (grid, grid, grid)
    L = grid * h
    top = 0.8 * (grid / 512.0) + h / 2 if grid >= 64 else 4 * h + h / 2
    j0 = int(np.floor(top / h)) + 1
    margin = max(4, grid // 32)
    foot = grid - 2 * margin
    thick = int(np.ceil(n / 8.0 / (foot * foot)))
    lo, hi = [margin, j0, margin], [grid - margin, min(j0 + thick, grid - 4), grid - margin]
    per_plane = (hi[1] - lo[1]) * (hi[2] - lo[2]) * 8          # candidates per i-plane of cells (no rejection in a box)
    n_keep = n
    if i_range is not None:
        # the global scene keeps the first n candidates in i-major order; a slab keeps its share of exactly those
        first = max(lo[0], i_range[0])
        n_keep = int(np.clip(n - (first - lo[0]) * per_plane, 0, None))
        lo[0], hi[0] = first, min(hi[0], i_range[1])
    pos = box_region(lo, hi, dims, h, seed, n_keep) if hi[0] > lo[0] and n_keep > 0 else np.zeros((0, 3), np.float32)
    g = 9.8
    a = np.deg2rad(tilt_deg)
    return _scene(pos, (0.0, 0.0, 0.0), dims, h, dt, [ground_collider(top, dims, h)], name=f"snow_slab_{grid}",
                  gravity=(np.float32(g * np.sin(a)), np.float32(-g * np.cos(a)), np.float32(0.0)), n_target=n)


def snow_slab_layer_counts(grid, n):
    """Particles of snow_slab(grid, n) per particle-block layer l (cells c with (c-1)>>2 == l), in closed form."""
    full = snow_slab_geometry(grid, n)
    lo, hi, per_plane = full
    counts = np.zeros((grid + 3) // 4, np.int64)
    left = n
    for c in range(lo[0], hi[0]):
        k = min(per_plane, left)
        if k <= 0:
            break
        counts[(c - 1) >> 2] += k
        left -= k
    return counts


def snow_slab_geometry(grid, n, h=0.05):
    top = 0.8 * (grid / 512.0) + h / 2 if grid >= 64 else 4 * h + h / 2
    j0 = int(np.floor(top / h)) + 1
    margin = max(4, grid // 32)
    foot = grid - 2 * margin
    thick = int(np.ceil(n / 8.0 / (foot * foot)))
    lo, hi = [margin, j0, margin], [grid - margin, min(j0 + thick, grid - 4), grid - margin]
    return lo, hi, (hi[1] - lo[1]) * (hi[2] - lo[2]) * 8


def small_ball(grid=32, radius_cells=5.0, h=0.05, dt=1e-5, seed=SEED, v0=(0.0, -200.0, 0.0), with_ground=True):
    """Small parity scene the CPU oracle finishes in seconds."""
    dims = (grid, grid, grid)
    L = grid * h
    top = 4 * h + h / 2
    center = (0.5 * L, top + (radius_cells + 2.5) * h, 0.5 * L)
    pos = ball(center, radius_cells * h, dims, h, seed)
    return _scene(pos, v0, dims, h, dt, [ground_collider(top, dims, h)] if with_ground else [], name=f"small_ball_{grid}")
