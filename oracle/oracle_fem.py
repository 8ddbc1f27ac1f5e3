"""ctypes wrapper of oracle/life_oracle_fem.c — the C restatement of the reference's structural solver (FEMBodyClass::dynamicFEM,
predictor, relaxation update).  TEST INFRASTRUCTURE, groundwork for SURVEY.md §8f row 3: nothing in the product
uses it.  Bodies are created from the description the compiled reference built (RefCase.fem_body)."""
import ctypes as C
import glob
import os
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblife_oracle_fem.so")
_lib = None


def lapack_path():
    """The OpenBLAS the compiled reference links (oracle/Makefile OPENBLAS): the same dgetrf_/dgetrs_ give the same bits."""
    p = glob.glob(os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs", "libopenblasp-*.so"))
    return p[0] if p else None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "life_oracle_fem.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        L = C.CDLL(LIB)
        L.orc_fem_create.restype = C.c_void_p
        L.orc_fem_create.argtypes = [C.c_int] * 3 + [C.c_void_p] * 10
        L.orc_fem_destroy.argtypes = [C.c_void_p]
        L.orc_fem_set_state.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_fem_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_fem_dynamic.argtypes = [C.c_void_p] * 6
        L.orc_fem_predict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_fem_relax.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_fem_bind_lapack.argtypes = [C.c_char_p]
        path = lapack_path()
        if path is None or L.orc_fem_bind_lapack(path.encode()) != 0:
            raise RuntimeError("oracle_fem: no LAPACK (dgetrf_/dgetrs_) to bind: " + str(path))
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class FemBody:
    """One flexible body.  `d` = RefCase.fem_body(fb)."""

    def __init__(self, d):
        self.L = load()
        self.n_dof, self.n_ibm = d["n_dof"], d["n_ibm"]
        consts = np.array([d["alpha"], d["delta"], d["Dt"], d["Dm"], d["gravityX"], d["gravityY"], d["ref_L"]])
        keep = [np.ascontiguousarray(d[k], np.float64) for k in ("pos0", "angle0", "el", "pm_zeta", "fm_z1", "fm_z2")]
        ints = [np.ascontiguousarray(d[k], np.int32) for k in ("pm_el", "fm_first", "fm_node")]
        pos0, angle0, el, pm_zeta, fm_z1, fm_z2 = keep
        pm_el, fm_first, fm_node = ints
        self.h = self.L.orc_fem_create(d["n_nodes"], d["n_bc"], d["n_ibm"], _p(consts), _p(pos0), _p(angle0), _p(el), _p(pm_el),
                                       _p(pm_zeta), _p(fm_first), _p(fm_node), _p(fm_z1), _p(fm_z2))

    def close(self):
        if self.h:
            self.L.orc_fem_destroy(self.h)
            self.h = None

    def set_state(self, st):
        st = np.ascontiguousarray(st, np.float64)
        assert st.shape == (11, self.n_dof)
        self.L.orc_fem_set_state(self.h, _p(st))

    def get_state(self):
        st = np.zeros((11, self.n_dof))
        self.L.orc_fem_get_state(self.h, _p(st))
        return st

    def dynamic(self, force, epsilon):
        """dynamicFEM -> (marker pos, marker vel, (subRes, subNum, subDen, resNR, itNR))"""
        force, epsilon = np.ascontiguousarray(force, np.float64), np.ascontiguousarray(epsilon, np.float64)
        pos, vel, res = np.zeros((self.n_ibm, 2)), np.zeros((self.n_ibm, 2)), np.zeros(5)
        rc = self.L.orc_fem_dynamic(self.h, _p(force), _p(epsilon), _p(pos), _p(vel), _p(res))
        assert rc == 0
        return pos, vel, tuple(res[:4]) + (int(res[4]),)

    def predict(self, t):
        pos, vel = np.zeros((self.n_ibm, 2)), np.zeros((self.n_ibm, 2))
        self.L.orc_fem_predict(self.h, int(t), _p(pos), _p(vel))
        return pos, vel

    def relax(self, relax):
        pos, vel = np.zeros((self.n_ibm, 2)), np.zeros((self.n_ibm, 2))
        self.L.orc_fem_relax(self.h, float(relax), _p(pos), _p(vel))
        return pos, vel
