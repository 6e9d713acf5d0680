"""ctypes binding of libmpm_b200.so (include/mpm_b200.h) — the host-side mirror of the reference's
MaterialPointMethod::LagrangeEulerView stage interface (material_point_method.hpp:174-233).

There is no CPU path: if the CUDA library is missing or no device is present, calls raise MpmError.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPM_B200_LIB", os.path.join(HERE, "libmpm_b200.so"))   # override: A/B builds during tuning
MIGRATE_FLOATS = 44


class MpmError(RuntimeError):
    pass


class MpmParams(C.Structure):
    _fields_ = [("h", C.c_float), ("youngs_modulus", C.c_float), ("poisson_ratio", C.c_float),
                ("hardening_xi", C.c_float), ("theta_c", C.c_float), ("theta_s", C.c_float),
                ("gravity", C.c_float * 3), ("friction_mu", C.c_float), ("p2g_variant", C.c_int),
                ("g2p_variant", C.c_int), ("fupdate_exact", C.c_int), ("stencil", C.c_int), ("reserved", C.c_int * 4)]


class MpmBoxCollider(C.Structure):
    _fields_ = [("world_to_local", C.c_float * 16), ("half_extent", C.c_float * 3), ("velocity", C.c_float * 3)]


class MpmBoxTransform(C.Structure):
    _fields_ = [("scale", C.c_float * 3), ("rotation_wxyz", C.c_float * 4), ("translation", C.c_float * 3),
                ("velocity", C.c_float * 3)]


class MpmImplicitParams(C.Structure):
    _fields_ = [("mu0", C.c_float), ("lambda0", C.c_float), ("xi", C.c_float), ("hardening", C.c_int), ("max_iters", C.c_int),
                ("ls_decrease", C.c_float), ("ls_tau", C.c_float), ("ls_max_iters", C.c_int), ("tol_grad", C.c_float),
                ("tol_step", C.c_float), ("reserved", C.c_int * 4)]


class MpmImplicitStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("evaluations", C.c_int), ("energy_start", C.c_double), ("energy_end", C.c_double),
                ("grad_norm_end", C.c_double), ("reserved", C.c_int * 4)]


class MpmStats(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("n_out_of_grid", C.c_int64), ("n_active_nodes", C.c_int64),
                ("n_particle_blocks", C.c_int64), ("n_grid_blocks", C.c_int64), ("substeps_done", C.c_int64),
                ("kernel_launches", C.c_int64), ("last_ms", C.c_float * 8), ("svd_failed", C.c_int32),
                ("reserved", C.c_int32 * 7)]


EXPORTS = ["mpm_default_params", "mpm_last_error", "mpm_device_count", "mpm_create", "mpm_create_slab", "mpm_destroy",
           "mpm_set_stream", "mpm_set_params", "mpm_upload_particles_aos", "mpm_upload_particles_soa",
           "mpm_download_particles_aos", "mpm_download_particles_soa", "mpm_download_render_buffers",
           "mpm_rasterize_particles_to_grid", "mpm_compute_particle_volumes_and_densities",
           "mpm_compute_explicit_grid_forces", "mpm_grid_velocities_update", "mpm_grid_based_collisions",
           "mpm_update_deformation_gradient", "mpm_update_particle_velocities", "mpm_update_particle_positions",
           "mpm_substep", "mpm_write_render_buffers_device", "mpm_download_grid", "mpm_upload_grid", "mpm_download_binning", "mpm_get_stats",
           "mpm_synchronize", "mpm_halo_bytes", "mpm_halo_pack", "mpm_halo_add", "mpm_substep_begin",
           "mpm_substep_end", "mpm_migrate_outgoing", "mpm_migrate_append", "mpm_set_pid_base", "mpm_download_live_particles", "mpm_migrate_buffer_bytes", "mpm_migrate_pack",
           "mpm_migrate_append_packed", "mpm_sync_counts", "mpm_set_migrate_capacity", "mpm_download_render_buffers_async",
           "mpm_wait_render_buffers", "mpm_box_collider_from_transform", "mpm_box_transform_move",
           "mpm_box_transform_flip_velocity", "mpm_fill_ball", "mpm_peer_export", "mpm_peer_connect", "mpm_peer_connect_ptr",
           "mpm_grid_device_ptr", "mpm_substep_begin_peer", "mpm_peer_export_migration", "mpm_peer_connect_migration",
           "mpm_peer_connect_migration_ptr", "mpm_migrate_peer", "mpm_reduce_invariants", "mpm_debug_p2g_profile", "mpm_load_obj", "mpm_free", "mpm_fill_mesh", "mpm_sphere_collider",
           "mpm_default_implicit_params", "mpm_time_integration", "mpm_energy", "mpm_energy_gradient"]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpmError(f"{LIB_PATH} is missing: build it with `python realtime-deformations_b200/build.py` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    if hasattr(L, "mpm_emulated_build") and os.environ.get("MPM_B200_ALLOW_EMULATION") != "1":
        # tests/emu builds a host emulation of the kernels for logic checks; it must never stand in for the CUDA library
        raise MpmError(f"{LIB_PATH} is the host-emulation test build, not the CUDA library (there is no CPU fallback)")
    fp, vp, i64 = C.POINTER(C.c_float), C.c_void_p, C.c_int64
    sz = C.c_size_t
    L.mpm_default_params.argtypes = [C.POINTER(MpmParams)]
    L.mpm_last_error.restype = C.c_char_p
    L.mpm_create.argtypes = [C.POINTER(MpmParams), C.c_int, C.c_int, C.c_int, i64, C.POINTER(vp)]
    L.mpm_create_slab.argtypes = [C.POINTER(MpmParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64, i64, C.POINTER(vp)]
    L.mpm_destroy.argtypes = [vp]
    L.mpm_set_stream.argtypes = [vp, vp]
    L.mpm_set_params.argtypes = [vp, C.POINTER(MpmParams)]
    L.mpm_upload_particles_aos.argtypes = [vp, vp, i64] + [sz] * 8
    L.mpm_download_particles_aos.argtypes = [vp, vp, i64] + [sz] * 8
    L.mpm_upload_particles_soa.argtypes = [vp, i64] + [fp] * 7
    L.mpm_download_particles_soa.argtypes = [vp, i64] + [fp] * 7
    L.mpm_download_render_buffers.argtypes = [vp, i64, vp, vp, C.c_float]
    L.mpm_download_render_buffers_async.argtypes = [vp, i64, vp, C.c_float]
    L.mpm_write_render_buffers_device.argtypes = [vp, i64, vp, vp, C.c_float]
    L.mpm_wait_render_buffers.argtypes = [vp]
    for n in ("mpm_rasterize_particles_to_grid", "mpm_compute_particle_volumes_and_densities",
              "mpm_compute_explicit_grid_forces", "mpm_update_particle_velocities", "mpm_synchronize"):
        getattr(L, n).argtypes = [vp]
    for n in ("mpm_grid_velocities_update", "mpm_update_deformation_gradient", "mpm_update_particle_positions",
              "mpm_substep_begin"):
        getattr(L, n).argtypes = [vp, C.c_float]
    L.mpm_grid_based_collisions.argtypes = [vp, C.c_float, C.POINTER(MpmBoxCollider), C.c_int]
    L.mpm_default_implicit_params.argtypes = [C.POINTER(MpmImplicitParams)]
    L.mpm_default_implicit_params.restype = None
    L.mpm_time_integration.argtypes = [vp, C.c_float, C.POINTER(MpmImplicitParams), C.POINTER(MpmImplicitStats)]
    L.mpm_energy.argtypes = [vp, C.c_float, C.POINTER(MpmImplicitParams), fp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mpm_energy_gradient.argtypes = [vp, C.c_float, C.POINTER(MpmImplicitParams), fp, C.c_int, fp]
    L.mpm_substep_end.argtypes = [vp, C.c_float, C.POINTER(MpmBoxCollider), C.c_int]
    L.mpm_substep.argtypes = [vp, C.c_float, C.POINTER(MpmBoxCollider), C.c_int, C.c_int]
    L.mpm_download_grid.argtypes = [vp, fp]
    L.mpm_upload_grid.argtypes = [vp, fp]
    L.mpm_download_binning.argtypes = [vp, i64, vp, vp, vp]
    L.mpm_get_stats.argtypes = [vp, C.POINTER(MpmStats)]
    L.mpm_halo_bytes.argtypes = [vp]
    L.mpm_halo_bytes.restype = sz
    L.mpm_halo_pack.argtypes = [vp, C.c_int, vp]
    L.mpm_halo_add.argtypes = [vp, C.c_int, vp]
    L.mpm_migrate_outgoing.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(vp), C.POINTER(vp)]
    L.mpm_migrate_append.argtypes = [vp, vp, i64]
    L.mpm_set_migrate_capacity.argtypes = [vp, i64]
    L.mpm_migrate_buffer_bytes.argtypes = [vp]
    L.mpm_migrate_buffer_bytes.restype = sz
    L.mpm_migrate_pack.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.mpm_migrate_append_packed.argtypes = [vp, vp]
    L.mpm_sync_counts.argtypes = [vp]
    L.mpm_set_pid_base.argtypes = [vp, i64]
    L.mpm_download_live_particles.argtypes = [vp, i64, C.POINTER(i64), fp, vp]
    L.mpm_load_obj.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(i64)]
    L.mpm_free.argtypes = [vp]
    L.mpm_free.restype = None
    L.mpm_fill_mesh.argtypes = [C.POINTER(C.c_float), i64, C.c_float, vp, vp, fp, i64, C.POINTER(i64), C.POINTER(i64)]
    L.mpm_sphere_collider.argtypes = [fp, C.c_float, fp, C.POINTER(MpmBoxCollider)]
    L.mpm_debug_p2g_profile.argtypes = [vp, C.POINTER(C.c_int64), C.c_int]
    L.mpm_reduce_invariants.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.mpm_box_collider_from_transform.argtypes = [C.POINTER(MpmBoxTransform), C.POINTER(MpmBoxCollider)]
    L.mpm_box_transform_move.argtypes = [C.POINTER(MpmBoxTransform), C.c_float]
    L.mpm_box_transform_flip_velocity.argtypes = [C.POINTER(MpmBoxTransform)]
    L.mpm_peer_export.argtypes = [vp, vp]
    L.mpm_peer_connect.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.mpm_peer_connect_ptr.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.mpm_grid_device_ptr.argtypes = [vp, C.POINTER(vp)]
    L.mpm_substep_begin_peer.argtypes = [vp, C.c_float, C.c_int]
    L.mpm_peer_export_migration.argtypes = [vp, vp, vp]
    L.mpm_peer_connect_migration.argtypes = [vp, vp, vp]
    L.mpm_peer_connect_migration_ptr.argtypes = [vp, vp, vp]
    L.mpm_migrate_peer.argtypes = [vp, C.c_int]
    L.mpm_fill_ball.argtypes = [fp, C.c_float, C.c_float, vp, vp, fp, i64, C.POINTER(i64), C.POINTER(i64)]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise MpmError(f"libmpm_b200 error {rc}: {lib().mpm_last_error().decode()}")


def default_params(**kw):
    p = MpmParams()
    lib().mpm_default_params(C.byref(p))
    known = {f[0] for f in MpmParams._fields_}
    for k, v in kw.items():
        if k not in known:          # a ctypes Structure would silently grow a Python attribute instead
            raise TypeError(f"unknown parameter {k!r}; fields are {sorted(known)}")
        if k == "gravity":
            p.gravity[:] = [float(x) for x in v]
        elif k in ("p2g_variant", "g2p_variant", "fupdate_exact", "stencil"):
            setattr(p, k, int(v))
        else:
            setattr(p, k, float(v))
    return p


def default_implicit_params(**kw):
    q = MpmImplicitParams()
    lib().mpm_default_implicit_params(C.byref(q))
    known = {f[0] for f in MpmImplicitParams._fields_}
    for k, v in kw.items():
        if k not in known:
            raise TypeError(f"unknown implicit parameter {k!r}; fields are {sorted(known)}")
        setattr(q, k, int(v) if k in ("hardening", "max_iters", "ls_max_iters") else float(v))
    return q


def make_colliders(w2l, half, vel=None):
    w2l = np.asarray(w2l, np.float32).reshape(-1, 16)
    half = np.asarray(half, np.float32).reshape(-1, 3)
    nc = w2l.shape[0]
    vel = np.zeros((nc, 3), np.float32) if vel is None else np.asarray(vel, np.float32).reshape(-1, 3)
    arr = (MpmBoxCollider * max(nc, 1))()
    for i in range(nc):
        arr[i].world_to_local[:] = w2l[i].tolist()
        arr[i].half_extent[:] = half[i].tolist()
        arr[i].velocity[:] = vel[i].tolist()
    return arr, nc


def box_transform(scale, rotation_wxyz, translation, velocity=(0.0, 0.0, 0.0)):
    """The pose a reference MeshCollider's sdf reads (hpp:80-83) + its velocity, as the C ABI's MpmBoxTransform."""
    t = MpmBoxTransform()
    t.scale[:] = [float(np.float32(x)) for x in scale]
    t.rotation_wxyz[:] = [float(np.float32(x)) for x in rotation_wxyz]
    t.translation[:] = [float(np.float32(x)) for x in translation]
    t.velocity[:] = [float(np.float32(x)) for x in velocity]
    return t


RAND_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


def fill_ball(origin, radius, h, capacity, rnd=None):
    """LagrangeEulerView::initializeParticles' fill rule (cpp:18-63) through the C ABI: (positions (n, 3) in acceptance
    order, candidates that did not fit). rnd: Python callable with libc rand()'s contract, None = libc rand() itself."""
    o = np.ascontiguousarray(origin, np.float32)
    pos = np.zeros((max(int(capacity), 1), 3), np.float32)
    n, miss = C.c_int64(0), C.c_int64(0)
    cb = RAND_FN(lambda _user: int(rnd())) if rnd is not None else None
    _ck(lib().mpm_fill_ball(_fp(o), float(radius), float(h), C.cast(cb, C.c_void_p) if cb is not None else None, None, _fp(pos),
                            int(capacity), C.byref(n), C.byref(miss)))
    return pos[:n.value].copy(), miss.value


def colliders_from_transforms(transforms):
    """MpmBoxCollider array from MpmBoxTransform poses through the library's own glm-order arithmetic."""
    nc = len(transforms)
    arr = (MpmBoxCollider * max(nc, 1))()
    for i, t in enumerate(transforms):
        _ck(lib().mpm_box_collider_from_transform(C.byref(t), C.byref(arr[i])))
    return arr, nc


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def load_obj(path):
    """Triangles (n, 3, 3) of a Wavefront OBJ through mpm_load_obj."""
    ptr, n = C.POINTER(C.c_float)(), C.c_int64()
    _ck(lib().mpm_load_obj(str(path).encode(), C.byref(ptr), C.byref(n)))
    tri = np.ctypeslib.as_array(ptr, shape=(n.value * 9,)).copy().reshape(n.value, 3, 3)
    lib().mpm_free(ptr)
    return tri


def fill_mesh(tri, h, capacity):
    """Particle positions inside a closed triangle mesh: initializeParticles' fill rule (libc rand()) through mpm_fill_mesh."""
    tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
    pos = np.empty((capacity, 3), np.float32)
    nw, nm = C.c_int64(), C.c_int64()
    _ck(lib().mpm_fill_mesh(_fp(tri), tri.shape[0], float(h), None, None, _fp(pos), capacity, C.byref(nw), C.byref(nm)))
    return pos[:nw.value].copy(), nm.value


def sphere_collider(centre, radius, velocity=(0.0, 0.0, 0.0)):
    c = MpmBoxCollider()
    _ck(lib().mpm_sphere_collider(_fp(np.asarray(centre, np.float32)), float(radius), _fp(np.asarray(velocity, np.float32)), C.byref(c)))
    return c


def _f32(a, shape):
    if a is None:
        return None
    a = np.ascontiguousarray(a, np.float32)
    assert a.size == int(np.prod(shape)), (a.shape, shape)
    return a


class Sim:
    """Mirror of LagrangeEulerView (hpp:174-233): same stage names, same call order, state resident in HBM."""

    def __init__(self, max_i, max_j, max_k, n_particles, params=None, slab=None, capacity=None):
        L = lib()
        self.L = L
        self.MAX_I, self.MAX_J, self.MAX_K = max_i, max_j, max_k
        self.n = int(n_particles)
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        if slab is None:
            _ck(L.mpm_create(C.byref(self.params), max_i, max_j, max_k, self.n, C.byref(h)))
        else:
            cap = int(capacity if capacity is not None else n_particles)
            _ck(L.mpm_create_slab(C.byref(self.params), max_i, max_j, max_k, slab[0], slab[1], self.n, cap, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.mpm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- particles -------------------------------------------------------------------------------
    def upload(self, pos, vel, mass, volume=None, FE=None, FP=None, B=None):
        n = self.n
        pos = _f32(pos, (n, 3)); vel = _f32(vel, (n, 3))
        mass = _f32(np.broadcast_to(np.asarray(mass, np.float32), (n,)), (n,))
        volume = _f32(volume, (n,)); FE = _f32(FE, (n, 9)); FP = _f32(FP, (n, 9)); B = _f32(B, (n, 9))
        _ck(self.L.mpm_upload_particles_soa(self.h, n, _fp(pos), _fp(vel), _fp(mass), _fp(volume), _fp(FE), _fp(FP), _fp(B)))

    def upload_state35(self, s):
        """oracle/ref_driver layout: mass, vel[3], volume, pos[3], FE[9], FP[9], B[9] — also exercises the AoS entry
        point with the strides/offsets a struct Particle binding would pass."""
        s = np.ascontiguousarray(s, np.float32).reshape(self.n, 35)
        _ck(self.L.mpm_upload_particles_aos(self.h, s.ctypes.data, self.n, 140, 0, 4, 16, 20, 32, 68, 104))

    def download_state35(self):
        s = np.zeros((self.n, 35), np.float32)
        _ck(self.L.mpm_download_particles_aos(self.h, s.ctypes.data, self.n, 140, 0, 4, 16, 20, 32, 68, 104))
        return s

    def download(self):
        n = self.n
        out = dict(pos=np.empty((n, 3), np.float32), vel=np.empty((n, 3), np.float32), mass=np.empty(n, np.float32),
                   volume=np.empty(n, np.float32), FE=np.empty((n, 9), np.float32), FP=np.empty((n, 9), np.float32),
                   B=np.empty((n, 9), np.float32))
        _ck(self.L.mpm_download_particles_soa(self.h, n, _fp(out["pos"]), _fp(out["vel"]), _fp(out["mass"]),
                                              _fp(out["volume"]), _fp(out["FE"]), _fp(out["FP"]), _fp(out["B"])))
        return out

    def render_buffers(self, size=0.02, xyzs=None, rgba=None):
        xyzs = np.empty((self.n, 4), np.float32) if xyzs is None else xyzs
        rgba = np.empty((self.n, 4), np.uint8) if rgba is None else rgba
        _ck(self.L.mpm_download_render_buffers(self.h, self.n, xyzs.ctypes.data, rgba.ctypes.data, size))
        return xyzs, rgba

    def capacity_rows(self):
        """Rows a render buffer needs in slab mode (storage order incl. retired slots): the current slot bound."""
        return int(self.stats().n_particles)

    def render_buffers_async(self, xyzs_pinned_ptr, n, size=0.02):
        _ck(self.L.mpm_download_render_buffers_async(self.h, int(n), C.c_void_p(xyzs_pinned_ptr), size))

    def write_render_buffers_device(self, d_xyzs_ptr, d_rgba_ptr=None, n=None, size=0.02):
        """Instance buffers straight into caller-owned device memory (mapped GL buffers, torch tensors ...)."""
        _ck(self.L.mpm_write_render_buffers_device(self.h, int(self.n if n is None else n), C.c_void_p(d_xyzs_ptr),
                                                   C.c_void_p(d_rgba_ptr) if d_rgba_ptr else None, size))

    def wait_render_buffers(self):
        _ck(self.L.mpm_wait_render_buffers(self.h))

    # ---- reference stages (main.cpp:192-218) ---------------------------------------------------------
    def rasterizeParticlesToGrid(self):
        _ck(self.L.mpm_rasterize_particles_to_grid(self.h))

    def computeParticleVolumesAndDensities(self):
        _ck(self.L.mpm_compute_particle_volumes_and_densities(self.h))

    def computeExplicitGridForces(self):
        _ck(self.L.mpm_compute_explicit_grid_forces(self.h))

    def gridVelocitiesUpdate(self, dt):
        _ck(self.L.mpm_grid_velocities_update(self.h, dt))

    def gridBasedCollisions(self, dt, colliders, nc):
        _ck(self.L.mpm_grid_based_collisions(self.h, dt, colliders, nc))

    def updateDeformationGradient(self, dt):
        _ck(self.L.mpm_update_deformation_gradient(self.h, dt))

    def updateParticleVelocities(self):
        _ck(self.L.mpm_update_particle_velocities(self.h))

    def updateParticlePositions(self, dt):
        _ck(self.L.mpm_update_particle_positions(self.h, dt))

    # ---- implicit time integration (LagrangeEulerView::timeIntegration, cpp:211-233) ----
    def timeIntegration(self, dt, params=None):
        q = params if params is not None else default_implicit_params()
        st = MpmImplicitStats()
        _ck(self.L.mpm_time_integration(self.h, dt, C.byref(q), C.byref(st)))
        return st

    def energy(self, dt, trial_velocity=None, relative=False, params=None):
        """(Energy, ElasticPotential) of cpp:160-185 at a dense (I*J*K, 3) trial velocity field (None: the grid's own)."""
        q = params if params is not None else default_implicit_params()
        tv = None if trial_velocity is None else _f32(trial_velocity, (self.MAX_I * self.MAX_J * self.MAX_K, 3))
        e, el = C.c_double(0), C.c_double(0)
        _ck(self.L.mpm_energy(self.h, dt, C.byref(q), None if tv is None else _fp(tv), int(relative), C.byref(e), C.byref(el)))
        return e.value, el.value

    def energy_gradient(self, dt, trial_velocity=None, relative=False, params=None):
        q = params if params is not None else default_implicit_params()
        n = self.MAX_I * self.MAX_J * self.MAX_K
        tv = None if trial_velocity is None else _f32(trial_velocity, (n, 3))
        g = np.empty((n, 3), np.float32)
        _ck(self.L.mpm_energy_gradient(self.h, dt, C.byref(q), None if tv is None else _fp(tv), int(relative), _fp(g)))
        return g

    def staged_substep(self, dt, colliders, nc):
        self.rasterizeParticlesToGrid(); self.computeExplicitGridForces(); self.gridVelocitiesUpdate(dt)
        self.gridBasedCollisions(dt, colliders, nc); self.updateDeformationGradient(dt)
        self.updateParticleVelocities(); self.updateParticlePositions(dt)

    def substep(self, dt, colliders, nc, n=1):
        _ck(self.L.mpm_substep(self.h, dt, colliders, nc, n))

    def substep_begin(self, dt):
        _ck(self.L.mpm_substep_begin(self.h, dt))

    def substep_end(self, dt, colliders, nc):
        _ck(self.L.mpm_substep_end(self.h, dt, colliders, nc))

    # ---- peer-memory halo (include/mpm_b200.h) -------------------------------------------
    def peer_export(self):
        h = (C.c_ubyte * 64)()
        _ck(self.L.mpm_peer_export(self.h, C.cast(h, C.c_void_p)))
        return bytes(h)

    def peer_connect(self, lower_handle, lower_layers, upper_handle, upper_layers):
        lo = (C.c_ubyte * 64).from_buffer_copy(lower_handle) if lower_handle is not None else None
        up = (C.c_ubyte * 64).from_buffer_copy(upper_handle) if upper_handle is not None else None
        _ck(self.L.mpm_peer_connect(self.h, C.cast(lo, C.c_void_p) if lo is not None else None, int(lower_layers),
                                    C.cast(up, C.c_void_p) if up is not None else None, int(upper_layers)))

    def peer_connect_ptr(self, lower_grid, lower_layers, upper_grid, upper_layers):
        _ck(self.L.mpm_peer_connect_ptr(self.h, C.c_void_p(lower_grid) if lower_grid else None, int(lower_layers),
                                        C.c_void_p(upper_grid) if upper_grid else None, int(upper_layers)))

    def grid_device_ptr(self):
        p = C.c_void_p()
        _ck(self.L.mpm_grid_device_ptr(self.h, C.byref(p)))
        return p.value

    def substep_begin_peer(self, dt, phase):
        _ck(self.L.mpm_substep_begin_peer(self.h, dt, int(phase)))

    def peer_export_migration(self):
        d, u = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
        _ck(self.L.mpm_peer_export_migration(self.h, C.cast(d, C.c_void_p), C.cast(u, C.c_void_p)))
        return bytes(d), bytes(u)

    def peer_connect_migration(self, lower_up_handle, upper_down_handle):
        lo = (C.c_ubyte * 64).from_buffer_copy(lower_up_handle) if lower_up_handle is not None else None
        up = (C.c_ubyte * 64).from_buffer_copy(upper_down_handle) if upper_down_handle is not None else None
        _ck(self.L.mpm_peer_connect_migration(self.h, C.cast(lo, C.c_void_p) if lo is not None else None,
                                              C.cast(up, C.c_void_p) if up is not None else None))

    def peer_connect_migration_ptr(self, lower_up_buf, upper_down_buf):
        _ck(self.L.mpm_peer_connect_migration_ptr(self.h, C.c_void_p(lower_up_buf) if lower_up_buf else None,
                                                  C.c_void_p(upper_down_buf) if upper_down_buf else None))

    def migrate_peer(self, phase):
        _ck(self.L.mpm_migrate_peer(self.h, int(phase)))

    # ---- diagnostics --------------------------------------------------------------------------------
    def grid(self):
        g = np.empty((self.MAX_I * self.MAX_J * self.MAX_K, 7), np.float32)
        _ck(self.L.mpm_download_grid(self.h, _fp(g)))
        return g

    def set_grid(self, g):
        g = np.ascontiguousarray(g, np.float32).reshape(-1, 7)
        _ck(self.L.mpm_upload_grid(self.h, _fp(g)))

    def binning(self):
        cells = np.empty((self.n, 3), np.int32); key = np.empty(self.n, np.int32); ids = np.empty(self.n, np.int32)
        _ck(self.L.mpm_download_binning(self.h, self.n, cells.ctypes.data, key.ctypes.data, ids.ctypes.data))
        return cells, key, ids

    def stats(self):
        st = MpmStats()
        _ck(self.L.mpm_get_stats(self.h, C.byref(st)))
        return st

    def synchronize(self):
        _ck(self.L.mpm_synchronize(self.h))

    # ---- slab decomposition plumbing (device pointers; the exchange itself is the caller's, see multi.py) ----
    def halo_bytes(self):
        return int(self.L.mpm_halo_bytes(self.h))

    def halo_pack(self, upper, dev_ptr):
        _ck(self.L.mpm_halo_pack(self.h, int(upper), C.c_void_p(dev_ptr)))

    def halo_add(self, upper, dev_ptr):
        _ck(self.L.mpm_halo_add(self.h, int(upper), C.c_void_p(dev_ptr)))

    def migrate_outgoing(self):
        nd, nu, pd, pu = C.c_int64(), C.c_int64(), C.c_void_p(), C.c_void_p()
        _ck(self.L.mpm_migrate_outgoing(self.h, C.byref(nd), C.byref(nu), C.byref(pd), C.byref(pu)))
        return nd.value, nu.value, pd.value, pu.value

    def migrate_append(self, dev_ptr, n):
        _ck(self.L.mpm_migrate_append(self.h, C.c_void_p(dev_ptr), int(n)))

    def set_migrate_capacity(self, records):
        _ck(self.L.mpm_set_migrate_capacity(self.h, int(records)))

    def migrate_buffer_bytes(self):
        return int(self.L.mpm_migrate_buffer_bytes(self.h))

    def migrate_pack(self):
        pd, pu = C.c_void_p(), C.c_void_p()
        _ck(self.L.mpm_migrate_pack(self.h, C.byref(pd), C.byref(pu)))
        return pd.value, pu.value

    def migrate_append_packed(self, dev_ptr):
        _ck(self.L.mpm_migrate_append_packed(self.h, C.c_void_p(dev_ptr)))

    def sync_counts(self):
        _ck(self.L.mpm_sync_counts(self.h))

    def set_pid_base(self, base):
        _ck(self.L.mpm_set_pid_base(self.h, int(base)))

    def download_live(self, capacity):
        st = np.empty((capacity, 35), np.float32); pid = np.empty(capacity, np.int32); n = C.c_int64()
        _ck(self.L.mpm_download_live_particles(self.h, capacity, C.byref(n), _fp(st), pid.ctypes.data))
        return st[:n.value], pid[:n.value]

    def invariants(self):
        """{count, id_sum, id_hash} (exact integers mod 2^64) and {mass, momentum[3], mass_y} (fp64 sums) of the live particles."""
        f = (C.c_double * 5)(); i = (C.c_uint64 * 3)()
        _ck(self.L.mpm_reduce_invariants(self.h, f, i))
        return {"count": int(i[0]), "id_sum": int(i[1]), "id_hash": int(i[2]), "mass": float(f[0]),
                "momentum": [float(f[1]), float(f[2]), float(f[3])], "mass_y": float(f[4])}

    def set_params(self, params):
        self.params = params
        _ck(self.L.mpm_set_params(self.h, C.byref(params)))

    def set_stream(self, cuda_stream):
        _ck(self.L.mpm_set_stream(self.h, C.c_void_p(cuda_stream)))
