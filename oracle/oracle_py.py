"""TEST INFRASTRUCTURE — ctypes binding of oracle/libmpm_oracle.so (the CPU restatement, mpm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The library is built on demand with the pinned flags of oracle/Makefile (gcc is on every box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmpm_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ref_mpm")

STATE_W = 35  # mass, vel[3], volume, pos[3], FE[9], FP[9], B[9]
COL = dict(mass=slice(0, 1), vel=slice(1, 4), volume=slice(4, 5), pos=slice(5, 8),
           FE=slice(8, 17), FP=slice(17, 26), B=slice(26, 35))


class OracleParams(C.Structure):
    _fields_ = [("h", C.c_float), ("E", C.c_float), ("nu", C.c_float), ("xi", C.c_float),
                ("theta_c", C.c_float), ("theta_s", C.c_float), ("gravity", C.c_float * 3),
                ("friction", C.c_float), ("stencil", C.c_int)]


class ImplicitParams(C.Structure):
    """OracleImplicitParams (mpm_oracle.h): the objective's constants and the vendored optimiser's settings."""
    _fields_ = [("mu0", C.c_float), ("lambda0", C.c_float), ("xi", C.c_float), ("hardening", C.c_int), ("max_iters", C.c_int),
                ("ls_decrease", C.c_float), ("ls_tau", C.c_float), ("ls_max_iters", C.c_int), ("tol_grad", C.c_float),
                ("tol_step", C.c_float), ("gradient", C.c_int)]


OBJECTIVE_FN = C.CFUNCTYPE(C.c_float, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int)


class BoxCollider(C.Structure):
    _fields_ = [("world_to_local", C.c_float * 16), ("half_extent", C.c_float * 3), ("velocity", C.c_float * 3)]


LIB_FMA_PATH = os.path.join(HERE, "libmpm_oracle_fma.so")


def build(force=False, fma=False):
    src = os.path.join(HERE, "mpm_oracle.c")
    path = LIB_FMA_PATH if fma else LIB_PATH
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, os.path.basename(path)])
    return path


_lib = None
_lib_fma = None


def lib(fma=False):
    """fma=True: the same restatement compiled with FMA contraction — only for measuring a scene's noise floor."""
    global _lib, _lib_fma
    if fma:
        if _lib_fma is None:
            _lib_fma = _bind(C.CDLL(build(fma=True)))
        return _lib_fma
    if _lib is None:
        _lib = _bind(C.CDLL(build()))
    return _lib


def _bind(L):
    if True:
        fp = C.POINTER(C.c_float)
        L.oracle_default_params.argtypes = [C.POINTER(OracleParams)]
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int] * 4 + [C.POINTER(OracleParams)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_particles.argtypes = [C.c_void_p, fp]
        L.oracle_get_particles.argtypes = [C.c_void_p, fp]
        L.oracle_get_grid.argtypes = [C.c_void_p, fp]
        L.oracle_set_grid.argtypes = [C.c_void_p, fp]
        L.oracle_num_used_cells.argtypes = [C.c_void_p]
        L.oracle_num_out_of_grid.argtypes = [C.c_void_p]
        L.oracle_cell_indices.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        for name in ("oracle_rasterize_particles_to_grid", "oracle_compute_particle_volumes_and_densities",
                     "oracle_compute_explicit_grid_forces", "oracle_update_particle_velocities"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_grid_velocities_update.argtypes = [C.c_void_p, C.c_float]
        L.oracle_grid_based_collisions.argtypes = [C.c_void_p, C.c_float, C.POINTER(BoxCollider), C.c_int]
        L.oracle_update_deformation_gradient.argtypes = [C.c_void_p, C.c_float]
        L.oracle_update_deformation_gradient.restype = C.c_int
        L.oracle_update_particle_positions.argtypes = [C.c_void_p, C.c_float]
        L.oracle_substep.argtypes = [C.c_void_p, C.c_float, C.POINTER(BoxCollider), C.c_int, C.c_int]
        L.oracle_weight.argtypes = [C.c_float]
        L.oracle_weight.restype = C.c_float
        L.oracle_svd3.argtypes = [fp, fp, fp, fp]
        L.oracle_svd3.restype = C.c_int
        L.oracle_polar_rotation.argtypes = [fp, fp]
        L.oracle_box_sdf.argtypes = [C.POINTER(BoxCollider), fp]
        L.oracle_box_sdf.restype = C.c_float
        L.oracle_body_collision.argtypes = [fp, fp, C.POINTER(BoxCollider), C.c_int, C.c_float, fp]
        ip = C.POINTER(ImplicitParams)
        L.oracle_default_implicit_params.argtypes = [ip]
        L.oracle_weight_derivative.argtypes = [C.c_float]
        L.oracle_weight_derivative.restype = C.c_float
        L.oracle_used_cells.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.oracle_energy.argtypes = [C.c_void_p, fp, C.c_float, ip]
        L.oracle_energy.restype = C.c_float
        L.oracle_energy_gradient.argtypes = [C.c_void_p, fp, C.c_float, ip, C.POINTER(C.c_double)]
        L.oracle_lbfgs.argtypes = [C.c_void_p, fp, C.c_int, C.c_float, ip, OBJECTIVE_FN, C.c_void_p, C.POINTER(C.c_int)]
        L.oracle_lbfgs.restype = C.c_int
        L.oracle_time_integration.argtypes = [C.c_void_p, C.c_float, ip, C.POINTER(C.c_int)]
        L.oracle_time_integration.restype = C.c_int
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_params(**kw):
    p = OracleParams()
    lib().oracle_default_params(C.byref(p))
    known = {f[0] for f in OracleParams._fields_}
    for k, v in kw.items():
        if k not in known:          # a ctypes Structure would silently grow a Python attribute instead
            raise TypeError(f"unknown parameter {k!r}; fields are {sorted(known)}")
        if k == "gravity":
            p.gravity[:] = [float(x) for x in v]
        elif k == "stencil":
            p.stencil = int(v)
        else:
            setattr(p, k, float(v))
    return p


def make_colliders(w2l, half, vel=None):
    """w2l: (nc,16) glm column-major inverse transforms; half: (nc,3); vel: (nc,3) or None."""
    w2l = np.asarray(w2l, np.float32).reshape(-1, 16)
    half = np.asarray(half, np.float32).reshape(-1, 3)
    nc = w2l.shape[0]
    vel = np.zeros((nc, 3), np.float32) if vel is None else np.asarray(vel, np.float32).reshape(-1, 3)
    arr = (BoxCollider * max(nc, 1))()
    for i in range(nc):
        arr[i].world_to_local[:] = w2l[i].tolist()
        arr[i].half_extent[:] = half[i].tolist()
        arr[i].velocity[:] = vel[i].tolist()
    return arr, nc


def colliders_from_ref_dump(raw):
    """oracle/ref_driver.cpp colliders.f32 rows: scale[3], quat wxyz[4], translation[3], velocity[3], inverse mat4[16]."""
    raw = np.asarray(raw, np.float32).reshape(-1, 29)
    return make_colliders(raw[:, 13:29], raw[:, 0:3], raw[:, 10:13])


class Oracle:
    def __init__(self, I, J, K, n, params=None, threads=1, fma=False):
        self.L = lib(fma)
        self.I, self.J, self.K, self.n = I, J, K, n
        self.params = params if params is not None else default_params()
        self.h = self.L.oracle_create(I, J, K, n, C.byref(self.params))
        self.L.oracle_set_threads(self.h, threads)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float32).reshape(self.n, STATE_W)
        self.L.oracle_set_particles(self.h, _fp(s))

    def state(self):
        s = np.empty((self.n, STATE_W), np.float32)
        self.L.oracle_get_particles(self.h, _fp(s))
        return s

    def grid(self):
        g = np.empty((self.I * self.J * self.K, 7), np.float32)
        self.L.oracle_get_grid(self.h, _fp(g))
        return g

    def set_grid(self, g):
        g = np.ascontiguousarray(g, np.float32).reshape(self.I * self.J * self.K, 7)
        self.L.oracle_set_grid(self.h, _fp(g))

    def cells(self):
        c = np.empty((self.n, 3), np.int32)
        self.L.oracle_cell_indices(self.h, c.ctypes.data_as(C.POINTER(C.c_int)))
        return c

    def num_used(self):
        return self.L.oracle_num_used_cells(self.h)

    def num_out_of_grid(self):
        return self.L.oracle_num_out_of_grid(self.h)

    def rasterize(self):
        self.L.oracle_rasterize_particles_to_grid(self.h)

    def volumes(self):
        self.L.oracle_compute_particle_volumes_and_densities(self.h)

    def forces(self):
        self.L.oracle_compute_explicit_grid_forces(self.h)

    def grid_velocities(self, dt):
        self.L.oracle_grid_velocities_update(self.h, dt)

    def collisions(self, dt, colliders, nc):
        self.L.oracle_grid_based_collisions(self.h, dt, colliders, nc)

    def fupdate(self, dt):
        return self.L.oracle_update_deformation_gradient(self.h, dt)

    def g2p(self):
        self.L.oracle_update_particle_velocities(self.h)

    def advect(self, dt):
        self.L.oracle_update_particle_positions(self.h, dt)

    def substep(self, dt, colliders, nc, nsteps=1):
        self.L.oracle_substep(self.h, dt, colliders, nc, nsteps)

    # ---- implicit time integration (cpp:160-233; dead code in the reference) ----
    def used_cells(self):
        u = np.empty(self.num_used(), np.int32)
        self.L.oracle_used_cells(self.h, u.ctypes.data_as(C.POINTER(C.c_int)))
        return u

    def energy(self, vel_used, dt, q=None):
        q = q if q is not None else default_implicit_params()
        v = np.ascontiguousarray(vel_used, np.float32).reshape(-1)
        assert v.size == 3 * self.num_used()
        return float(self.L.oracle_energy(self.h, _fp(v), dt, C.byref(q)))

    def energy_gradient(self, vel_used, dt, q=None):
        q = q if q is not None else default_implicit_params()
        v = np.ascontiguousarray(vel_used, np.float32).reshape(-1)
        g = np.empty(v.size, np.float64)
        self.L.oracle_energy_gradient(self.h, _fp(v), dt, C.byref(q), g.ctypes.data_as(C.POINTER(C.c_double)))
        return g.reshape(-1, 3)

    def time_integration(self, dt, q=None):
        q = q if q is not None else default_implicit_params()
        evals = C.c_int(0)
        it = self.L.oracle_time_integration(self.h, dt, C.byref(q), C.byref(evals))
        return it, evals.value


def default_implicit_params(**kw):
    q = ImplicitParams()
    lib().oracle_default_implicit_params(C.byref(q))
    for k, v in kw.items():
        if k not in {f[0] for f in ImplicitParams._fields_}:
            raise TypeError(f"unknown implicit parameter {k!r}")
        setattr(q, k, v)
    return q


def lbfgs(x0, fn, params=None):
    """The optimiser restatement on a caller-supplied objective fn(x) -> (value, grad): (iterations, x, evaluations)."""
    q = params if params is not None else default_implicit_params()
    x = np.ascontiguousarray(x0, np.float32).copy()

    def cb(_user, xp, gp, n):
        xv = np.ctypeslib.as_array(xp, shape=(n,))
        val, grad = fn(xv)
        if gp:
            np.ctypeslib.as_array(gp, shape=(n,))[:] = grad
        return float(val)
    evals = C.c_int(0)
    it = lib().oracle_lbfgs(None, _fp(x), x.size, 0.0, C.byref(q), OBJECTIVE_FN(cb), None, C.byref(evals))
    return it, x, evals.value


def initial_state(pos, vel, mass):
    """35-float state rows with FE = FP = I, B = 0, volume = 0 (set by volumes())."""
    n = pos.shape[0]
    s = np.zeros((n, STATE_W), np.float32)
    s[:, COL["mass"]] = np.asarray(mass, np.float32).reshape(-1, 1) if np.ndim(mass) else mass
    s[:, COL["vel"]] = vel
    s[:, COL["pos"]] = pos
    s[:, 8] = s[:, 12] = s[:, 16] = 1.0
    s[:, 17] = s[:, 21] = s[:, 25] = 1.0
    return s


def svd3(A):
    A = np.ascontiguousarray(A, np.float32).reshape(9)
    U = np.empty(9, np.float32); S = np.empty(3, np.float32); V = np.empty(9, np.float32)
    rc = lib().oracle_svd3(_fp(A), _fp(U), _fp(S), _fp(V))
    return rc, U.reshape(3, 3), S, V.reshape(3, 3)


def polar_rotation(F):
    F = np.ascontiguousarray(F, np.float32).reshape(9)
    R = np.empty(9, np.float32)
    lib().oracle_polar_rotation(_fp(F), _fp(R))
    return R


def have_reference_binary():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)
