"""ctypes wrapper around oracle/_ref/libref_<case>.so (the compiled, unmodified reference + oracle/ref_harness.cpp).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import
this.  One process can hold ONE reference case at a time per shared object (the reference keeps its case in
compile-time constants and the harness keeps one GridClass/ObjectsClass pair), so `RefCase` is a singleton per
case name; run different cases in different processes or sequentially with `.close()`.
"""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# LAPACK inside the reference must run single threaded: set before the library is mapped.
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available(case):
    return os.path.exists(os.path.join(REF_DIR, "libref_%s.so" % case))


class RefCase:
    """The reference program for one compile-time case, driven stage by stage."""

    FLAG_CM, FLAG_ORDERED, FLAG_UNI_EPS, FLAG_WOMERSLEY, FLAG_RAMP, FLAG_IBM, FLAG_FLEX = 1, 2, 4, 8, 16, 32, 64

    def __init__(self, case, quiet=True):
        path = os.path.join(REF_DIR, "libref_%s.so" % case)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle ref`)")
        self.case = case
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        self._cwd = os.getcwd()
        self.workdir = tempfile.mkdtemp(prefix="life_ref_%s_" % case)
        geo = os.path.join(REF_DIR, "input_%s" % case, "geometry.config")
        if os.path.exists(geo):
            os.makedirs(os.path.join(self.workdir, "input"))
            shutil.copy(geo, os.path.join(self.workdir, "input", "geometry.config"))
        L = self.lib
        L.ref_create.argtypes = [C.c_char_p]
        L.ref_get_subres.restype = C.c_double
        L.ref_get_relax.restype = C.c_double
        L.ref_ref_pressure.restype = C.c_double
        # the reference chats on stdout while constructing; silence it at the fd level
        saved = None
        if quiet:
            import sys
            sys.stdout.flush()
            saved = os.dup(1)
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
            os.close(devnull)
        try:
            rc = L.ref_create(self.workdir.encode())
        finally:
            if saved is not None:
                # flush the C++ side before restoring
                C.CDLL(None).fflush(None)
                os.dup2(saved, 1)
                os.close(saved)
            os.chdir(self._cwd)
        if rc != 0:
            raise RuntimeError("ref_create failed rc=%d" % rc)
        d = np.zeros(5, np.int32)
        L.ref_dims(d.ctypes.data_as(C.c_void_p))
        self.Nx, self.Ny, self.n_markers, self.n_bodies, self.n_bc = (int(x) for x in d)
        w = np.zeros(4, np.int32)
        L.ref_walls(w.ctypes.data_as(C.c_void_p))
        self.walls = tuple(int(x) for x in w)  # left, right, bottom, top
        self.flags = int(L.ref_flags())
        s = np.zeros(19, np.float64)
        L.ref_scalars(s.ctypes.data_as(C.c_void_p))
        (self.omega, self.Dx, self.Dt, self.Dm, self.Drho, self.inlet_ramp, self.womersley, self.dpdx, self.dpdy,
         self.gravityX, self.gravityY, self.height_p, self.nu_p, self.rho_p, self.subTol, self.uxInlet_p,
         self.uyInlet_p, self.ux0_p, self.uy0_p) = (float(x) for x in s)
        self.profile = int(L.ref_profile())
        self.central_moments = bool(self.flags & self.FLAG_CM)
        self.ordered = bool(self.flags & self.FLAG_ORDERED)
        self.has_ibm = bool(self.flags & self.FLAG_IBM)
        self.has_flex = bool(self.flags & self.FLAG_FLEX)
        self.ref_P = float(L.ref_ref_pressure())

    # ---- helpers -----------------------------------------------------------------------------------------------
    def _in_workdir(self, fn, *a):
        os.chdir(self.workdir)
        try:
            return fn(*a)
        finally:
            os.chdir(self._cwd)

    def close(self):
        if self.lib is not None:
            self.lib.ref_destroy()
            self.lib = None
            shutil.rmtree(self.workdir, ignore_errors=True)

    def _get(self, name, shape, dtype=np.float64):
        out = np.zeros(shape, dtype)
        getattr(self.lib, "ref_get_" + name)(out.ctypes.data_as(C.c_void_p))
        return out

    def _set(self, name, arr):
        arr = np.ascontiguousarray(arr, np.float64)
        getattr(self.lib, "ref_set_" + name)(arr.ctypes.data_as(C.c_void_p))

    # ---- time loop ---------------------------------------------------------------------------------------------
    @property
    def t(self):
        return int(self.lib.ref_get_t())

    @t.setter
    def t(self, v):
        self.lib.ref_set_t(int(v))

    def step(self, n=1):
        self._in_workdir(self.lib.ref_step, int(n))

    def lbm_kernel(self):
        self.lib.ref_lbm_kernel()

    def object_kernel(self):
        self._in_workdir(self.lib.ref_object_kernel)

    def recompute_object_vals(self):
        self.lib.ref_recompute_object_vals()

    def ibm_interp(self):
        self.lib.ref_ibm_interp()

    def fem_kernel(self):
        self.lib.ref_fem_kernel()

    def ibm_spread(self):
        self.lib.ref_ibm_spread()

    @property
    def subit(self):
        return int(self.lib.ref_get_subit())

    @subit.setter
    def subit(self, v):
        self.lib.ref_set_subit(int(v))

    @property
    def subres(self):
        return float(self.lib.ref_get_subres())

    @property
    def relax(self):
        """ObjectsClass::relax, the Aitken relaxation factor of the current sub-iteration (src/Objects.cpp:178-188)"""
        return float(self.lib.ref_get_relax())

    # ---- lattice state (reference layout: id = i*Ny + j) ----------------------------------------------------------
    def f(self):
        return self._get("f", (self.Nx, self.Ny, 9))

    def f_n(self):
        return self._get("f_n", (self.Nx, self.Ny, 9))

    def rho(self):
        return self._get("rho", (self.Nx, self.Ny))

    def rho_n(self):
        return self._get("rho_n", (self.Nx, self.Ny))

    def u(self):
        return self._get("u", (self.Nx, self.Ny, 2))

    def u_n(self):
        return self._get("u_n", (self.Nx, self.Ny, 2))

    def force_xy(self):
        return self._get("force_xy", (self.Nx, self.Ny, 2))

    def force_ibm(self):
        return self._get("force_ibm", (self.Nx, self.Ny, 2))

    def u_in(self):
        return self._get("u_in", (self.Ny, 2))

    def rho_in(self):
        return self._get("rho_in", (self.Ny,))

    def type(self):
        return self._get("type", (self.Nx, self.Ny), np.int32)

    def bcvec(self):
        return self._get("bcvec", (self.n_bc,), np.int32)

    def delu(self):
        out = np.zeros((self.Ny, 2))
        n = self.lib.ref_get_delu(out.ctypes.data_as(C.c_void_p))
        return out if n else None

    def set_state(self, f=None, rho=None, u=None, force_ibm=None, force_xy=None):
        for name, arr in (("f", f), ("rho", rho), ("u", u), ("force_ibm", force_ibm), ("force_xy", force_xy)):
            if arr is not None:
                self._set(name, arr)

    # ---- markers ---------------------------------------------------------------------------------------------------
    def markers(self):
        n = self.n_markers
        pos, vel, force = np.zeros((n, 2)), np.zeros((n, 2)), np.zeros((n, 2))
        ds, eps = np.zeros(n), np.zeros(n)
        body, flex = np.zeros(n, np.int32), np.zeros(n, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.ref_get_markers(p(pos), p(vel), p(force), p(ds), p(eps), p(body), p(flex))
        return dict(pos=pos, vel=vel, force=force, ds=ds, epsilon=eps, body=body, flex=flex)

    def set_marker_force(self, force):
        force = np.ascontiguousarray(force, np.float64)
        self.lib.ref_set_marker_force(force.ctypes.data_as(C.c_void_p))

    def set_marker_posvel(self, pos, vel):
        pos = np.ascontiguousarray(pos, np.float64)
        vel = np.ascontiguousarray(vel, np.float64)
        self.lib.ref_set_marker_posvel(pos.ctypes.data_as(C.c_void_p), vel.ctypes.data_as(C.c_void_p))

    def interp_values(self):
        n = self.n_markers
        rho, mom = np.zeros(n), np.zeros((n, 2))
        self.lib.ref_get_interp(rho.ctypes.data_as(C.c_void_p), mom.ctypes.data_as(C.c_void_p))
        return rho, mom

    def supports(self):
        n = self.n_markers
        count = np.zeros(n, np.int32)
        idx, jdx = np.zeros((n, 9), np.int32), np.zeros((n, 9), np.int32)
        dirac = np.zeros((n, 9))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.ref_get_supports(p(count), p(idx), p(jdx), p(dirac))
        return count, idx, jdx, dirac

    def refresh_supports(self, with_epsilon=True):
        self.lib.ref_refresh_supports(int(bool(with_epsilon)))

    def write_restart(self):
        self._in_workdir(self.lib.ref_write_restart)
        return os.path.join(self.workdir, "Results", "Restart")

    def write_vtk(self):
        """Results/VTK/Fluid.<t>.vti written by the reference's own GridClass::writeVTK; returns its path."""
        os.makedirs(os.path.join(self.workdir, "Results", "VTK"), exist_ok=True)   # only made when the case defines VTK
        self._in_workdir(self.lib.ref_write_vtk)
        return os.path.join(self.workdir, "Results", "VTK", "Fluid.%d.vti" % self.t)

    def read_restart(self):
        """GridClass::readRestart on <workdir>/Results/Restart/Fluid.restart; returns tOffset."""
        return self._in_workdir(self.lib.ref_read_restart)

    # ---- flexible bodies (for the FEM restatement, oracle/life_oracle_fem.c) ----------------------------------------------
    def fem_count(self):
        return int(self.lib.ref_fem_count())

    def fem_body(self, fb):
        """Description of the fb-th flexible body as the reference built it (geometry, maps, constants)."""
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        d = np.zeros(7, np.int32)
        self.lib.ref_fem_dims(int(fb), p(d))
        nn, ne, ndof, nbc, nibm, nmap, sim = (int(x) for x in d)
        c = np.zeros(8)
        self.lib.ref_fem_constants(p(c))
        pos0, angle0, el = np.zeros((nn, 2)), np.zeros(nn), np.zeros((ne, 5))
        self.lib.ref_fem_geometry(int(fb), p(pos0), p(angle0), p(el))
        pm_el, pm_zeta = np.zeros(nibm, np.int32), np.zeros(nibm)
        fm_first, fm_node = np.zeros(ne + 1, np.int32), np.zeros(nmap, np.int32)
        fm_z1, fm_z2, marker = np.zeros(nmap), np.zeros(nmap), np.zeros(nibm, np.int32)
        self.lib.ref_fem_maps(int(fb), p(pm_el), p(pm_zeta), p(fm_first), p(fm_node), p(fm_z1), p(fm_z2), p(marker))
        return dict(n_nodes=nn, n_el=ne, n_dof=ndof, n_bc=nbc, n_ibm=nibm, sim_dofs=sim, alpha=c[0], delta=c[1], Dt=c[2], Dm=c[3],
                    Dx=c[4], gravityX=c[5], gravityY=c[6], ref_L=c[7], pos0=pos0, angle0=angle0, el=el, pm_el=pm_el, pm_zeta=pm_zeta,
                    fm_first=fm_first, fm_node=fm_node, fm_z1=fm_z1, fm_z2=fm_z2, marker=marker)

    def fem_get_state(self, fb, n_dof):
        out = np.zeros((11, n_dof))     # U, Udot, Udotdot, U_n, Udot_n, Udotdot_n, U_km1, R_k, R_km1, U_nm1, U_nm2
        self.lib.ref_fem_get_state(int(fb), out.ctypes.data_as(C.c_void_p))
        return out

    def fem_set_state(self, fb, state):
        state = np.ascontiguousarray(state, np.float64)
        self.lib.ref_fem_set_state(int(fb), state.ctypes.data_as(C.c_void_p))

    def fem_dynamic(self, fb):
        """FEMBodyClass::dynamicFEM of one body -> (subRes, subNum, subDen, resNR, itNR)"""
        out = np.zeros(5)
        self.lib.ref_fem_dynamic(int(fb), out.ctypes.data_as(C.c_void_p))
        return tuple(out[:4]) + (int(out[4]),)
