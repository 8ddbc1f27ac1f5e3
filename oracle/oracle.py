"""ctypes wrapper around oracle/_build/liblife_oracle.so (oracle/life_oracle.c, the CPU restatement).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblife_oracle.so")

FLUID, WALL, VELOCITY, FREESLIP, PRESSURE, CONVECTIVE = range(6)
PROFILE_UNIFORM, PROFILE_PARABOLIC, PROFILE_SHEAR, PROFILE_BOUNDARYLAYER = -1, 0, 1, 2


class Params(C.Structure):
    """run-time image of inc/params.h (struct orc_params)"""
    _fields_ = [("Nx", C.c_int64), ("Ny", C.c_int64),
                ("central_moments", C.c_int32), ("ordered", C.c_int32), ("uni_epsilon", C.c_int32),
                ("profile", C.c_int32),
                ("wall_left", C.c_int32), ("wall_right", C.c_int32), ("wall_bottom", C.c_int32),
                ("wall_top", C.c_int32),
                ("inlet_ramp", C.c_double), ("womersley", C.c_double), ("omega", C.c_double),
                ("height_p", C.c_double), ("rho_p", C.c_double), ("nu_p", C.c_double),
                ("ux0_p", C.c_double), ("uy0_p", C.c_double),
                ("gravityX", C.c_double), ("gravityY", C.c_double), ("dpdx", C.c_double), ("dpdy", C.c_double),
                ("uxInlet_p", C.c_double), ("uyInlet_p", C.c_double)]

    def __init__(self, **kw):
        super().__init__()
        self.profile = PROFILE_UNIFORM
        self.inlet_ramp = -1.0
        self.womersley = -1.0
        self.omega = 1.0
        self.height_p = 1.0
        self.rho_p = 1.0
        self.nu_p = 0.01
        self.uxInlet_p = 1.0
        self.wall_left = self.wall_right = self.wall_bottom = self.wall_top = WALL
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    """Compile the restatement (gcc, seconds)."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
            os.path.getmtime(os.path.join(HERE, "life_oracle.c")), os.path.getmtime(os.path.join(HERE, "life_oracle.h"))):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Params)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_scalings.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_t.argtypes = [C.c_void_p]
        L.orc_set_t.argtypes = [C.c_void_p, C.c_int32]
        L.orc_lbm_kernel.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int32]
        L.orc_array.restype = C.POINTER(C.c_double)
        L.orc_array.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
        L.orc_types.restype = C.POINTER(C.c_int32)
        L.orc_types.argtypes = [C.c_void_p]
        L.orc_bc_count.restype = C.c_int64
        L.orc_bc_count.argtypes = [C.c_void_p]
        L.orc_bc_ids.restype = C.POINTER(C.c_int64)
        L.orc_bc_ids.argtypes = [C.c_void_p]
        L.orc_normal.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_stream_target.restype = C.c_int64
        L.orc_stream_target.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32]
        L.orc_set_markers.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4
        L.orc_get_marker_force.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_marker_force.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_interp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_find_support.argtypes = [C.c_void_p]
        L.orc_get_supports.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_compute_ds.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_compute_epsilon.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.orc_get_ds_eps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ibm_interp.argtypes = [C.c_void_p]
        L.orc_ibm_spread.argtypes = [C.c_void_p]
        L.orc_dirac_delta.restype = C.c_double
        L.orc_dirac_delta.argtypes = [C.c_double]
        L.orc_equilibrium.restype = C.c_double
        L.orc_equilibrium.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32]
        _lib = L
    return _lib


_ARR = dict(f=0, f_n=1, rho=2, rho_n=3, u=4, u_n=5, force_xy=6, force_ibm=7, u_in=8, rho_in=9, delU=10)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One lattice (+ optional markers) advanced by the C restatement."""

    def __init__(self, params):
        self.L = lib()
        self.params = params
        self.h = self.L.orc_create(C.byref(params))
        self.Nx, self.Ny = int(params.Nx), int(params.Ny)
        s = np.zeros(6)
        self.L.orc_scalings(self.h, _p(s))
        self.Dx, self.Dt, self.Dm, self.Drho, self.tau, self.nu = (float(x) for x in s)
        self.n_markers = 0

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- arrays (views into the C arrays: copy if you need to keep them) ----
    def view(self, name):
        n = C.c_int64()
        ptr = self.L.orc_array(self.h, _ARR[name], C.byref(n))
        a = np.ctypeslib.as_array(ptr, shape=(n.value,))
        shape = dict(f=(self.Nx, self.Ny, 9), f_n=(self.Nx, self.Ny, 9), rho=(self.Nx, self.Ny),
                     rho_n=(self.Nx, self.Ny), u=(self.Nx, self.Ny, 2), u_n=(self.Nx, self.Ny, 2),
                     force_xy=(self.Nx, self.Ny, 2), force_ibm=(self.Nx, self.Ny, 2), u_in=(self.Ny, 2),
                     rho_in=(self.Ny,), delU=(self.Ny, 2))[name]
        return a.reshape(shape)

    def get(self, name):
        return self.view(name).copy()

    def set(self, name, value):
        self.view(name)[...] = value

    def types(self):
        ptr = self.L.orc_types(self.h)
        return np.ctypeslib.as_array(ptr, shape=(self.Nx * self.Ny,)).reshape(self.Nx, self.Ny).copy()

    def bcvec(self):
        n = self.L.orc_bc_count(self.h)
        return np.ctypeslib.as_array(self.L.orc_bc_ids(self.h), shape=(n,)).copy()

    def normal(self, i, j):
        nx, ny = C.c_int32(), C.c_int32()
        d = self.L.orc_normal(self.h, i, j, C.byref(nx), C.byref(ny))
        return nx.value, ny.value, d

    def stream_target(self, i, j, v):
        return int(self.L.orc_stream_target(self.h, i, j, v))

    # ---- time loop ----
    @property
    def t(self):
        return int(self.L.orc_get_t(self.h))

    @t.setter
    def t(self, v):
        self.L.orc_set_t(self.h, int(v))

    def lbm_kernel(self):
        self.L.orc_lbm_kernel(self.h)

    def step(self, n=1):
        self.L.orc_step(self.h, int(n))

    # ---- markers ----
    def set_markers(self, pos=None, vel=None, ds=None, eps=None, n=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (pos, vel, ds, eps)]
        if n is None:
            n = len(arrs[0]) if arrs[0] is not None else self.n_markers
        self.n_markers = int(n)
        self.L.orc_set_markers(self.h, self.n_markers, *[_p(a) for a in arrs])

    def find_support(self):
        return int(self.L.orc_find_support(self.h))

    def supports(self):
        n = self.n_markers
        count = np.zeros(n, np.int32)
        idx, jdx = np.zeros((n, 9), np.int32), np.zeros((n, 9), np.int32)
        dirac = np.zeros((n, 9))
        self.L.orc_get_supports(self.h, _p(count), _p(idx), _p(jdx), _p(dirac))
        return count, idx, jdx, dirac

    def compute_ds(self, first, count):
        self.L.orc_compute_ds(self.h, first, count)

    def compute_epsilon(self, first, count):
        self.L.orc_compute_epsilon(self.h, first, count)

    def ds_eps(self):
        ds, eps = np.zeros(self.n_markers), np.zeros(self.n_markers)
        self.L.orc_get_ds_eps(self.h, _p(ds), _p(eps))
        return ds, eps

    def ibm_interp(self):
        self.L.orc_ibm_interp(self.h)

    def ibm_spread(self):
        self.L.orc_ibm_spread(self.h)

    def marker_force(self):
        f = np.zeros((self.n_markers, 2))
        self.L.orc_get_marker_force(self.h, _p(f))
        return f

    def set_marker_force(self, force):
        force = np.ascontiguousarray(force, np.float64)
        self.L.orc_set_marker_force(self.h, _p(force))

    def interp_values(self):
        rho, mom = np.zeros(self.n_markers), np.zeros((self.n_markers, 2))
        self.L.orc_get_interp(self.h, _p(rho), _p(mom))
        return rho, mom


def params_from_ref(ref):
    """orc_params equal to the compile-time case of a RefCase (oracle/refharness.py)."""
    return Params(Nx=ref.Nx, Ny=ref.Ny, central_moments=int(ref.central_moments), ordered=int(ref.ordered),
                  uni_epsilon=int(bool(ref.flags & ref.FLAG_UNI_EPS)), profile=ref.profile,
                  wall_left=ref.walls[0], wall_right=ref.walls[1], wall_bottom=ref.walls[2], wall_top=ref.walls[3],
                  inlet_ramp=ref.inlet_ramp, womersley=ref.womersley, omega=ref.omega, height_p=ref.height_p,
                  rho_p=ref.rho_p, nu_p=ref.nu_p, ux0_p=ref.ux0_p, uy0_p=ref.uy0_p, gravityX=ref.gravityX,
                  gravityY=ref.gravityY, dpdx=ref.dpdx, dpdy=ref.dpdy, uxInlet_p=ref.uxInlet_p,
                  uyInlet_p=ref.uyInlet_p)
