"""ctypes binding of the C ABI in include/life_b200.h (liblife_b200.so).

This is the thinnest possible layer: one Python method per exported function, numpy arrays in the reference's
layout in and out.  No arithmetic happens here and there is no fallback: if the shared library is missing or no
B200 is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblife_b200.so")

ABI_VERSION = 2
FLUID, WALL, VELOCITY, FREESLIP, PRESSURE, CONVECTIVE = range(6)
BGK, CENTRAL_MOMENTS = 0, 1
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_SHUFFLE, KERNEL_TMA, KERNEL_QUAD = 0, 1, 2, 3, 4
OK, E_ARG, E_CUDA, E_NCCL, E_STATE, E_SUPPORT, E_NOMEM, E_IO = range(8)
IO_SYNC, IO_ASYNC = 0, 1

# every symbol include/life_b200.h declares (tests check the library exports exactly these)
EXPORTS = [
    "life_abi_version", "life_nccl_unique_id", "life_create", "life_destroy", "life_last_error", "life_slab",
    "life_slab_range",
    "life_upload_state", "life_upload_begin", "life_upload_columns", "life_upload_end", "life_download_columns",
    "life_download_macro", "life_download_state", "life_max_speed", "life_step", "life_step_n",
    "life_sync", "life_ibm_set_markers", "life_ibm_interp", "life_ibm_spread", "life_ibm_compute_epsilon", "life_ibm_assemble_epsilon", "life_ibm_set_forces",
    "life_ibm_get_interp", "life_ibm_get_supports", "life_get_boundary", "life_get_types", "life_launch_count",
    "life_bulk_kernel_ms", "life_set_profiling", "life_membw",
    "life_vtk_frame", "life_write_vtk", "life_write_restart", "life_io_wait", "life_io_busy", "life_io_stats", "life_io_set_staging",
    "life_read_restart",
    "life_fem_create", "life_fem_set_state", "life_fem_get_state", "life_fem_predict", "life_fem_relax", "life_fem_dynamic",
    "life_ibm_get_markers", "life_fsi_move", "life_fsi_force", "life_ibm_set_epsilon", "life_ibm_get_marker_state",
]


class Config(C.Structure):
    """struct life_config"""
    _fields_ = [("abi_version", C.c_int32), ("collision", C.c_int32),
                ("Nx", C.c_int64), ("Ny", C.c_int64),
                ("omega", C.c_double),
                ("wall_left", C.c_int32), ("wall_right", C.c_int32), ("wall_bottom", C.c_int32),
                ("wall_top", C.c_int32),
                ("inlet_ramp", C.c_double),
                ("Dx", C.c_double), ("Dt", C.c_double), ("Dm", C.c_double), ("Drho", C.c_double),
                ("womersley", C.c_double), ("height_p", C.c_double), ("nu_p", C.c_double),
                ("gravity_x", C.c_double), ("gravity_y", C.c_double), ("dpdx", C.c_double), ("dpdy", C.c_double),
                ("ordered", C.c_int32), ("device", C.c_int32),
                ("stream", C.c_void_p),
                ("rank", C.c_int32), ("nranks", C.c_int32),
                ("nccl_id", C.c_void_p),
                ("kernel", C.c_int32), ("tune", C.c_int32), ("exact", C.c_int32), ("inplace", C.c_int32), ("reserved", C.c_int32 * 4)]

    def __init__(self, **kw):
        super().__init__()
        self.abi_version = ABI_VERSION
        self.device = -1
        self.inlet_ramp = -1.0
        self.womersley = -1.0
        self.Drho = 1.0
        self.wall_left = self.wall_right = self.wall_bottom = self.wall_top = WALL
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class FemBody(C.Structure):
    """struct life_fem_body"""
    _fields_ = [("n_nodes", C.c_int32), ("n_bc", C.c_int32), ("n_markers", C.c_int32),
                ("alpha", C.c_double), ("delta", C.c_double), ("gravity_x", C.c_double), ("gravity_y", C.c_double),
                ("ref_L", C.c_double),
                ("pos0", C.c_void_p), ("angle0", C.c_void_p), ("element", C.c_void_p),
                ("marker", C.c_void_p), ("marker_element", C.c_void_p), ("marker_zeta", C.c_void_p),
                ("map_first", C.c_void_p), ("map_marker", C.c_void_p), ("map_zeta1", C.c_void_p), ("map_zeta2", C.c_void_p)]


class LifeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("liblife_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """dlopen liblife_b200.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(LIB_PATH + " is missing: build it with `python -m life_b200.build` "
                                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.life_abi_version.restype = C.c_int
    L.life_nccl_unique_id.argtypes = [vp]
    L.life_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.life_destroy.argtypes = [vp]
    L.life_last_error.restype = C.c_char_p
    L.life_last_error.argtypes = [vp]
    L.life_slab.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.life_slab_range.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    L.life_upload_state.argtypes = [vp] + [vp] * 7
    L.life_upload_begin.argtypes = [vp, vp, vp]
    L.life_upload_columns.argtypes = [vp, i64, i64] + [vp] * 5
    L.life_upload_end.argtypes = [vp]
    L.life_download_columns.argtypes = [vp, i64, i64] + [vp] * 4
    L.life_download_macro.argtypes = [vp, vp, vp]
    L.life_download_state.argtypes = [vp, vp, vp, vp, vp]
    L.life_max_speed.argtypes = [vp, C.POINTER(dbl), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
    L.life_step.argtypes = [vp, i32]
    L.life_step_n.argtypes = [vp, i32, i32]
    L.life_sync.argtypes = [vp]
    L.life_ibm_set_markers.argtypes = [vp, i64, vp, vp, vp, vp]
    L.life_ibm_interp.argtypes = [vp, vp]
    L.life_ibm_spread.argtypes = [vp]
    L.life_ibm_compute_epsilon.argtypes = [vp, i64, vp, vp, vp]
    L.life_ibm_assemble_epsilon.argtypes = [vp, i64, vp, vp, vp]
    L.life_ibm_set_forces.argtypes = [vp, vp]
    L.life_ibm_get_interp.argtypes = [vp, vp, vp]
    L.life_ibm_get_supports.argtypes = [vp, vp, vp, vp, vp]
    L.life_get_boundary.argtypes = [vp, C.POINTER(i64), vp, vp, vp, vp, vp]
    L.life_get_types.argtypes = [vp, vp]
    L.life_launch_count.restype = i64
    L.life_launch_count.argtypes = [vp]
    L.life_bulk_kernel_ms.argtypes = [vp, C.POINTER(dbl), C.POINTER(i64)]
    L.life_set_profiling.argtypes = [vp, i32]
    L.life_vtk_frame.argtypes = [i64, i64, dbl, vp, i64, C.POINTER(i64), vp, i64, C.POINTER(i64)]
    L.life_write_vtk.argtypes = [vp, C.c_char_p, dbl, dbl, i32]
    L.life_write_restart.argtypes = [vp, C.c_char_p, i32, i32]
    L.life_io_wait.argtypes = [vp]
    L.life_io_busy.argtypes = [vp, C.POINTER(i32)]
    L.life_io_stats.argtypes = [vp, C.POINTER(dbl), C.POINTER(i64), C.POINTER(i32)]
    L.life_io_set_staging.argtypes = [vp, i64]
    L.life_read_restart.argtypes = [vp, C.c_char_p, vp, vp, vp, C.POINTER(i32)]
    L.life_fem_create.argtypes = [vp, i32, vp]
    L.life_fem_set_state.argtypes = [vp, i32, vp]
    L.life_fem_get_state.argtypes = [vp, i32, vp]
    L.life_fem_predict.argtypes = [vp, i32]
    L.life_fem_relax.argtypes = [vp, dbl]
    L.life_fem_dynamic.argtypes = [vp, vp, vp]
    L.life_ibm_get_markers.argtypes = [vp, vp, vp]
    L.life_fsi_move.argtypes = [vp, i32, i32, dbl]
    L.life_fsi_force.argtypes = [vp, vp, vp]
    L.life_ibm_set_epsilon.argtypes = [vp, vp]
    L.life_ibm_get_marker_state.argtypes = [vp, vp, vp, vp]
    _lib = L
    return L


def membw(mode, nbytes=8 << 30, iters=5, device=-1):
    """GB/s of the library's plain streaming kernels (life_membw): the memory-system ceiling next to the sweep's number."""
    L = load()
    L.life_membw.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_double)]
    out = C.c_double(0.0)
    rc = L.life_membw(int(device), int(mode), int(nbytes), int(iters), C.byref(out))
    if rc != OK:
        raise LifeError(rc, "life_membw(mode=%d)" % mode)
    return out.value


def nccl_unique_id():
    buf = (C.c_char * 128)()
    rc = load().life_nccl_unique_id(buf)
    if rc:
        raise LifeError(rc, load().life_last_error(None).decode())
    return bytes(buf)


def slab_range(Nx, nranks, rank):
    """Columns [begin, end) of rank `rank` (life_slab_range: pure arithmetic, no device needed)."""
    b, e = C.c_int64(), C.c_int64()
    rc = load().life_slab_range(int(Nx), int(nranks), int(rank), C.byref(b), C.byref(e))
    if rc:
        raise LifeError(rc, "life_slab_range: bad arguments")
    return b.value, e.value


def vtk_frame(Nx, Ny, Dx):
    """(head, tail) bytes of Fluid.<t>.vti around the raw arrays (life_vtk_frame: pure host arithmetic, no device needed)."""
    L = load()
    hl, tl = C.c_int64(), C.c_int64()
    rc = L.life_vtk_frame(int(Nx), int(Ny), float(Dx), None, 0, C.byref(hl), None, 0, C.byref(tl))
    if rc:
        raise LifeError(rc, "life_vtk_frame: bad arguments")
    head, tail = C.create_string_buffer(hl.value), C.create_string_buffer(tl.value)
    rc = L.life_vtk_frame(int(Nx), int(Ny), float(Dx), head, hl.value, C.byref(hl), tail, tl.value, C.byref(tl))
    if rc:
        raise LifeError(rc, "life_vtk_frame: bad arguments")
    return head.raw, tail.raw


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One life_ctx.  Methods map one-to-one onto the C entry points."""

    def __init__(self, cfg, nccl_id=None):
        self.L = load()
        self._id_buf = None
        if nccl_id is not None:
            self._id_buf = C.create_string_buffer(bytes(nccl_id), 128)
            cfg.nccl_id = C.cast(self._id_buf, C.c_void_p)
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.L.life_create(C.byref(cfg), C.byref(h))
        if rc:
            raise LifeError(rc, self.L.life_last_error(None).decode())
        self.h = h
        b, e = C.c_int64(), C.c_int64()
        self.L.life_slab(self.h, C.byref(b), C.byref(e))
        self.i_begin, self.i_end = b.value, e.value
        self.nxl = self.i_end - self.i_begin
        self.Ny = int(cfg.Ny)
        self.n_markers = 0

    def _ck(self, rc):
        if rc:
            raise LifeError(rc, self.L.life_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.life_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----
    def upload_state(self, f, rho=None, u=None, force_xy=None, force_ibm=None, u_in=None, rho_in=None):
        arrs = [_f64(a) for a in (f, rho, u, force_xy, force_ibm, u_in, rho_in)]
        n = self.nxl * self.Ny
        assert arrs[0].size == n * 9, "f has the wrong size for this slab"
        self._ck(self.L.life_upload_state(self.h, *[_ptr(a) for a in arrs]))

    def upload_begin(self, u_in=None, rho_in=None):
        u_in, rho_in = _f64(u_in), _f64(rho_in)
        self._ck(self.L.life_upload_begin(self.h, _ptr(u_in), _ptr(rho_in)))

    def upload_columns(self, il0, ncols, f, rho=None, u=None, force_xy=None, force_ibm=None):
        arrs = [_f64(a) for a in (f, rho, u, force_xy, force_ibm)]
        assert arrs[0].size == ncols * self.Ny * 9, "f has the wrong size for this column range"
        self._ck(self.L.life_upload_columns(self.h, int(il0), int(ncols), *[_ptr(a) for a in arrs]))

    def upload_end(self):
        self._ck(self.L.life_upload_end(self.h))

    def download_columns_into(self, il0, ncols, f=None, rho=None, u=None, force_ibm=None):
        self._ck(self.L.life_download_columns(self.h, int(il0), int(ncols), _ptr(f), _ptr(rho), _ptr(u), _ptr(force_ibm)))

    def download_macro(self, rho=True, u=True):
        r = np.empty((self.nxl, self.Ny)) if rho else None
        v = np.empty((self.nxl, self.Ny, 2)) if u else None
        self._ck(self.L.life_download_macro(self.h, _ptr(r), _ptr(v)))
        return r, v

    def download_state(self):
        f = np.empty((self.nxl, self.Ny, 9))
        r = np.empty((self.nxl, self.Ny))
        v = np.empty((self.nxl, self.Ny, 2))
        fi = np.empty((self.nxl, self.Ny, 2))
        self._ck(self.L.life_download_state(self.h, _ptr(f), _ptr(r), _ptr(v), _ptr(fi)))
        return dict(f=f, rho=r, u=v, force_ibm=fi)

    def download_into(self, f=None, rho=None, u=None, force_ibm=None):
        self._ck(self.L.life_download_state(self.h, _ptr(f), _ptr(rho), _ptr(u), _ptr(force_ibm)))

    def download_macro_into(self, rho=None, u=None):
        self._ck(self.L.life_download_macro(self.h, _ptr(rho), _ptr(u)))

    def max_speed(self):
        v, hn, ni, nj = C.c_double(), C.c_int32(), C.c_int64(), C.c_int64()
        self._ck(self.L.life_max_speed(self.h, C.byref(v), C.byref(hn), C.byref(ni), C.byref(nj)))
        return v.value, bool(hn.value), ni.value, nj.value

    # ---- stepping ----
    def step(self, t):
        self._ck(self.L.life_step(self.h, int(t)))

    def step_n(self, t_first, n):
        self._ck(self.L.life_step_n(self.h, int(t_first), int(n)))

    def sync(self):
        self._ck(self.L.life_sync(self.h))

    # ---- markers ----
    def ibm_set_markers(self, pos, vel, ds, eps):
        pos, vel, ds, eps = _f64(pos), _f64(vel), _f64(ds), _f64(eps)
        n = len(ds)
        self.n_markers = n
        self._ck(self.L.life_ibm_set_markers(self.h, n, _ptr(pos), _ptr(vel), _ptr(ds), _ptr(eps)))

    def ibm_interp(self):
        out = np.empty((self.n_markers, 2))
        self._ck(self.L.life_ibm_interp(self.h, _ptr(out)))
        return out

    def ibm_spread(self):
        self._ck(self.L.life_ibm_spread(self.h))

    def ibm_compute_epsilon(self, groups):
        """groups: list of integer index arrays (one per body); returns the epsilon of every marker"""
        first = np.zeros(len(groups) + 1, np.int64)
        first[1:] = np.cumsum([len(g) for g in groups])
        members = np.ascontiguousarray(np.concatenate(groups) if groups else np.zeros(0), np.int64)
        out = np.empty(self.n_markers)
        self._ck(self.L.life_ibm_compute_epsilon(self.h, len(groups), _ptr(first), _ptr(members), _ptr(out)))
        return out

    def ibm_assemble_epsilon(self, groups):
        """list of dim x dim matrices A (row-major, as handed to solveLAPACK), one per group"""
        first = np.zeros(len(groups) + 1, np.int64)
        first[1:] = np.cumsum([len(g) for g in groups])
        members = np.ascontiguousarray(np.concatenate(groups) if groups else np.zeros(0), np.int64)
        flat = np.empty(int(sum(len(g) ** 2 for g in groups)))
        self._ck(self.L.life_ibm_assemble_epsilon(self.h, len(groups), _ptr(first), _ptr(members), _ptr(flat)))
        out, off = [], 0
        for g in groups:
            out.append(flat[off:off + len(g) ** 2].reshape(len(g), len(g)).copy())
            off += len(g) ** 2
        return out

    def ibm_set_forces(self, force):
        force = _f64(force)
        self._ck(self.L.life_ibm_set_forces(self.h, _ptr(force)))

    def ibm_get_interp(self):
        r, m = np.empty(self.n_markers), np.empty((self.n_markers, 2))
        self._ck(self.L.life_ibm_get_interp(self.h, _ptr(r), _ptr(m)))
        return r, m

    def ibm_get_supports(self):
        n = self.n_markers
        count = np.zeros(n, np.int32)
        idx, jdx = np.zeros((n, 9), np.int32), np.zeros((n, 9), np.int32)
        dirac = np.zeros((n, 9))
        self._ck(self.L.life_ibm_get_supports(self.h, _ptr(count), _ptr(idx), _ptr(jdx), _ptr(dirac)))
        return count, idx, jdx, dirac

    # ---- device-fed files ----
    def write_vtk(self, path, rho_p, ref_P=0.0, mode=IO_SYNC):
        self._ck(self.L.life_write_vtk(self.h, os.fsencode(path), float(rho_p), float(ref_P), int(mode)))

    def write_restart(self, path, t, mode=IO_SYNC):
        self._ck(self.L.life_write_restart(self.h, os.fsencode(path), int(t), int(mode)))

    def io_wait(self):
        self._ck(self.L.life_io_wait(self.h))

    def io_busy(self):
        b = C.c_int32()
        self._ck(self.L.life_io_busy(self.h, C.byref(b)))
        return bool(b.value)

    def io_stats(self):
        s, b, a = C.c_double(), C.c_int64(), C.c_int32()
        self._ck(self.L.life_io_stats(self.h, C.byref(s), C.byref(b), C.byref(a)))
        return s.value, b.value, bool(a.value)

    def io_set_staging(self, nbytes):
        self._ck(self.L.life_io_set_staging(self.h, int(nbytes)))

    def read_restart(self, path, force_xy=None, u_in=None, rho_in=None):
        """Returns the time step stored in the file (GridClass::tOffset)."""
        fxy, u_in, rho_in = _f64(force_xy), _f64(u_in), _f64(rho_in)
        t = C.c_int32()
        self._ck(self.L.life_read_restart(self.h, os.fsencode(path), _ptr(fxy), _ptr(u_in), _ptr(rho_in), C.byref(t)))
        return t.value

    # ---- structural solver of the flexible bodies (not yet verified on a B200, see include/life_b200.h) ----
    def fem_create(self, bodies):
        """bodies: list of dicts with n_nodes, n_bc, alpha, delta, gravityX, gravityY, ref_L, pos0, angle0, el, marker, pm_el, pm_zeta,
        fm_first, fm_node, fm_z1, fm_z2 (one dict per body, keys as in struct life_fem_body: see tests/test_gpu_fem.py)"""
        arr = (FemBody * len(bodies))()
        self._fem_keep = []
        self._fem_dofs = []
        for k, d in enumerate(bodies):
            f64 = {n: np.ascontiguousarray(d[n], np.float64) for n in ("pos0", "angle0", "el", "pm_zeta", "fm_z1", "fm_z2")}
            i32 = {n: np.ascontiguousarray(d[n], np.int32) for n in ("marker", "pm_el", "fm_first", "fm_node")}
            self._fem_keep.append((f64, i32))
            b = arr[k]
            b.n_nodes, b.n_bc, b.n_markers = int(d["n_nodes"]), int(d["n_bc"]), len(i32["marker"])
            b.alpha, b.delta, b.gravity_x, b.gravity_y, b.ref_L = (float(d[n]) for n in ("alpha", "delta", "gravityX", "gravityY", "ref_L"))
            b.pos0, b.angle0, b.element = (f64[n].ctypes.data for n in ("pos0", "angle0", "el"))
            b.marker, b.marker_element, b.marker_zeta = i32["marker"].ctypes.data, i32["pm_el"].ctypes.data, f64["pm_zeta"].ctypes.data
            b.map_first, b.map_marker = i32["fm_first"].ctypes.data, i32["fm_node"].ctypes.data
            b.map_zeta1, b.map_zeta2 = f64["fm_z1"].ctypes.data, f64["fm_z2"].ctypes.data
            self._fem_dofs.append(3 * int(d["n_nodes"]))
        self._ck(self.L.life_fem_create(self.h, len(bodies), C.cast(arr, C.c_void_p)))

    def fem_set_state(self, body, state):
        state = np.ascontiguousarray(state, np.float64)
        assert state.shape == (11, self._fem_dofs[body])
        self._ck(self.L.life_fem_set_state(self.h, int(body), _ptr(state)))

    def fem_get_state(self, body):
        st = np.zeros((11, self._fem_dofs[body]))
        self._ck(self.L.life_fem_get_state(self.h, int(body), _ptr(st)))
        return st

    def fem_predict(self, t):
        self._ck(self.L.life_fem_predict(self.h, int(t)))

    def fem_relax(self, relax):
        self._ck(self.L.life_fem_relax(self.h, float(relax)))

    def fem_dynamic(self):
        """-> (sums [subRes, subNum, subDen], per-body array [n_bodies, 5] = subRes, subNum, subDen, resNR, itNR)"""
        sums, per = np.zeros(3), np.zeros((len(self._fem_dofs), 5))
        self._ck(self.L.life_fem_dynamic(self.h, _ptr(sums), _ptr(per)))
        return sums, per

    def fsi_move(self, t, sub_it, relax=0.0):
        self._ck(self.L.life_fsi_move(self.h, int(t), int(sub_it), C.c_double(relax)))

    def fsi_force(self):
        """interp + dynamicFEM on the device -> (sums, per-body array) like fem_dynamic"""
        sums, per = np.zeros(3), np.zeros((len(self._fem_dofs), 5))
        self._ck(self.L.life_fsi_force(self.h, _ptr(sums), _ptr(per)))
        return sums, per

    def ibm_set_epsilon(self, eps):
        eps = _f64(eps)
        assert eps.size == self.n_markers
        self._ck(self.L.life_ibm_set_epsilon(self.h, _ptr(eps)))

    def ibm_get_marker_state(self):
        force, ds, eps = np.zeros((self.n_markers, 2)), np.zeros(self.n_markers), np.zeros(self.n_markers)
        self._ck(self.L.life_ibm_get_marker_state(self.h, _ptr(force), _ptr(ds), _ptr(eps)))
        return force, ds, eps

    def ibm_get_markers(self):
        pos, vel = np.zeros((self.n_markers, 2)), np.zeros((self.n_markers, 2))
        self._ck(self.L.life_ibm_get_markers(self.h, _ptr(pos), _ptr(vel)))
        return pos, vel

    # ---- test / measurement hooks ----
    def boundary(self):
        n = C.c_int64()
        self._ck(self.L.life_get_boundary(self.h, C.byref(n), None, None, None, None, None))
        ids = np.zeros(n.value, np.int64)
        ty, nx, ny, nd = (np.zeros(n.value, np.int32) for _ in range(4))
        self._ck(self.L.life_get_boundary(self.h, C.byref(n), _ptr(ids), _ptr(ty), _ptr(nx), _ptr(ny), _ptr(nd)))
        return ids, ty, nx, ny, nd

    def types(self):
        t = np.zeros((self.nxl, self.Ny), np.int32)
        self._ck(self.L.life_get_types(self.h, _ptr(t)))
        return t

    def launch_count(self):
        return int(self.L.life_launch_count(self.h))

    def set_profiling(self, on):
        self._ck(self.L.life_set_profiling(self.h, int(bool(on))))

    def bulk_kernel_ms(self):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.L.life_bulk_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
