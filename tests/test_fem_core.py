"""The device structural solver (life_b200/csrc/fem_core.h, prepared for SURVEY.md §8f row 3 — not yet part of liblife_b200.so)
run SERIALLY on the CPU (a CTA of one thread, barriers compiled out; tests/native/fem_core_host.cpp) against the compiled,
unmodified reference inside live fluid-structure runs: every predictor / relaxed update / dynamicFEM call of every flexible body
is repeated from the reference's state and must agree to rounding — the core uses its own LU instead of LAPACK and sums element
contributions in a different order, so the bar here is rounding, not bit identity — 1e-11 of the body length for displacements and marker positions, 1e-8 for the
velocities / accelerations / residual sums derived from them through 1/Dt and 1/Dt^2 (the
bit-identical restatement is oracle/life_oracle_fem.c, tests/test_oracle_fem.py).  This checks the LOGIC of the device code; what
a one-thread CTA cannot check — barrier placement between cooperating threads — has to be verified on a B200.
"""
import os
import subprocess
import sys

import pytest

from oracle import refharness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "fem_core_host.cpp")
CORE = os.path.join(ROOT, "life_b200", "csrc", "fem_core.h")
LIB = os.path.join(ROOT, "tests", "native", "_build", "libfem_core_host.so")


@pytest.fixture(scope="module")
def host_core():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", LIB, SRC])
    return LIB


SCRIPT = r'''
import ctypes as C, sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
L = C.CDLL(%(lib)r)
L.femc_create.restype = C.c_void_p
L.femc_create.argtypes = [C.c_int] * 3 + [C.c_void_p] * 10
for f in (L.femc_set_state, L.femc_get_state):
    f.argtypes = [C.c_void_p, C.c_void_p]
L.femc_dynamic.argtypes = [C.c_void_p] * 6
L.femc_predict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
L.femc_relax.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
L.femc_compute_ds.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
p = lambda a: a.ctypes.data_as(C.c_void_p)
case, steps = %(case)r, %(steps)d
r = RefCase(case)
nb = r.fem_count()
desc = [r.fem_body(fb) for fb in range(nb)]
keep, core = [], []
for d in desc:
    consts = np.array([d["alpha"], d["delta"], d["Dt"], d["Dm"], d["gravityX"], d["gravityY"], d["ref_L"]])
    arrs = [consts] + [np.ascontiguousarray(d[k], np.float64) for k in ("pos0", "angle0", "el")] + [np.ascontiguousarray(d["pm_el"], np.int32),
            np.ascontiguousarray(d["pm_zeta"], np.float64), np.ascontiguousarray(d["fm_first"], np.int32), np.ascontiguousarray(d["fm_node"], np.int32),
            np.ascontiguousarray(d["fm_z1"], np.float64), np.ascontiguousarray(d["fm_z2"], np.float64)]
    keep.append(arrs)
    core.append(L.femc_create(d["n_nodes"], d["n_bc"], d["n_ibm"], *[p(a) for a in arrs]))
worst = {}
def close(a, b, what, scale_floor, tol=1e-11):
    a, b = np.asarray(a, float), np.asarray(b, float)
    err = float(np.abs(a - b).max() / max(np.abs(b).max(), scale_floor))
    worst[what] = max(worst.get(what, 0.0), err)
    assert err < tol, (what, err)
def state(h, n):
    st = np.zeros((11, n)); L.femc_get_state(h, p(st)); return st
calls = 0
for step in range(steps):
    r.t = r.t + 1
    r.lbm_kernel()
    r.subit = 0
    while True:
        before = [r.fem_get_state(fb, desc[fb]["n_dof"]) for fb in range(nb)]
        r.recompute_object_vals()
        m = r.markers()
        for fb in range(nb):
            d, ids = desc[fb], desc[fb]["marker"]
            L.femc_set_state(core[fb], p(np.ascontiguousarray(before[fb])))
            pos, vel = np.zeros((d["n_ibm"], 2)), np.zeros((d["n_ibm"], 2))
            if r.subit == 0:
                L.femc_predict(core[fb], r.t, p(pos), p(vel))
            else:
                L.femc_relax(core[fb], r.relax, p(pos), p(vel))
            ref_st = r.fem_get_state(fb, d["n_dof"])
            close(pos, m["pos"][ids], "predict/relax marker pos", d["ref_L"])
            ds = np.zeros(d["n_ibm"])                      # computeDs of the moved markers (src/IBMNode.cpp:182-204): bit for bit
            L.femc_compute_ds(core[fb], p(np.ascontiguousarray(m["pos"][ids])), d["Dx"], p(ds))
            assert np.array_equal(ds, m["ds"][ids]), ("ds", float(np.abs(ds - m["ds"][ids]).max()))
            close(vel, m["vel"][ids], "predict/relax marker vel", d["ref_L"] / d["Dt"] * 1e-3, 1e-8)
            got = state(core[fb], d["n_dof"])
            for k in (0, 3, 6, 9, 10):          # U, U_n, U_km1, U_nm1, U_nm2
                close(got[k], ref_st[k], "predict/relax displacement vectors", d["ref_L"])
            for k, s in ((1, 1.0 / d["Dt"]), (2, 1.0 / d["Dt"] ** 2), (4, 1.0 / d["Dt"]), (5, 1.0 / d["Dt"] ** 2)):
                close(got[k], ref_st[k], "predict/relax velocity / acceleration vectors", d["ref_L"] * s * 1e-3, 1e-8)
        r.ibm_interp()
        m = r.markers()
        for fb in range(nb):
            d, ids = desc[fb], desc[fb]["marker"]
            st = r.fem_get_state(fb, d["n_dof"])
            L.femc_set_state(core[fb], p(np.ascontiguousarray(st)))
            pos, vel, res = np.zeros((d["n_ibm"], 2)), np.zeros((d["n_ibm"], 2)), np.zeros(5)
            force, eps = np.ascontiguousarray(m["force"][ids]), np.ascontiguousarray(m["epsilon"][ids])
            L.femc_dynamic(core[fb], p(force), p(eps), p(pos), p(vel), p(res))
            ref_res = r.fem_dynamic(fb)
            after = r.fem_get_state(fb, d["n_dof"])
            m2 = r.markers()
            assert abs(int(res[4]) - ref_res[4]) <= 1, ("Newton-Raphson iterations", res, ref_res)
            got = state(core[fb], d["n_dof"])
            close(got[0], after[0], "U after dynamicFEM", d["ref_L"])
            close(got[1], after[1], "Udot after dynamicFEM", d["ref_L"] / d["Dt"] * 1e-3, 1e-8)
            close(got[2], after[2], "Udotdot after dynamicFEM", d["ref_L"] / d["Dt"] ** 2 * 1e-3, 1e-8)
            close(got[7], after[7], "R_k", d["ref_L"]); close(got[8], after[8], "R_km1", d["ref_L"])
            close(pos, m2["pos"][ids], "marker pos", d["ref_L"])
            close(vel, m2["vel"][ids], "marker vel", d["ref_L"] / d["Dt"] * 1e-3, 1e-8)
            close(res[:3], ref_res[:3], "subRes, subNum, subDen", d["ref_L"] ** 2 * 1e-6, 1e-8)
            r.fem_set_state(fb, st)
            calls += 1
        r.set_marker_posvel(m["pos"], m["vel"])
        r.fem_kernel()
        r.subit = r.subit + 1
        if not (r.subit < 20 and r.subres > r.subTol):
            break
    r.ibm_spread()
print("%%s: %%d bodies, %%d steps, %%d dynamicFEM calls; worst relative differences: %%s" %% (case, nb, steps, calls, {k: float("%%.1e" %% v) for k, v in worst.items()}))
r.close()
print("OK")
'''

CASES = [("TurekHron", 80), ("InvertedFlag", 25), ("PELskin", 12), ("Honami", 6)]


@pytest.mark.parametrize("case,steps", CASES, ids=[c for c, _ in CASES])
def test_serial_device_core_matches_the_reference(case, steps, host_core):
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built (make -C oracle ref)" % case)
    p = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, lib=host_core, case=case, steps=steps)], capture_output=True,
                       text=True, timeout=900, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-3000:] + p.stderr[-3000:]
    print(p.stdout.strip().splitlines()[-2])


# ---- barrier placement: the CTA as real threads under ThreadSanitizer --------------------------------------------------------------
RACE_SRC = os.path.join(ROOT, "tests", "native", "fem_core_race.cpp")
RACE_EXE = os.path.join(ROOT, "tests", "native", "_build", "fem_core_race")

RECORD = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
case, steps, fb, out = %(case)r, %(steps)d, %(fb)d, %(out)r
r = RefCase(case)
d = r.fem_body(fb)
ids, n = d["marker"], d["n_dof"]
calls = []
for step in range(steps):
    r.t = r.t + 1
    r.lbm_kernel()
    r.subit = 0
    while True:
        before = r.fem_get_state(fb, n)
        r.recompute_object_vals()
        m = r.markers()
        calls.append((1 if r.subit == 0 else 2, r.t, r.relax, before, m["force"][ids], m["epsilon"][ids]))
        r.ibm_interp()
        m = r.markers()
        calls.append((0, r.t, 0.0, r.fem_get_state(fb, n), m["force"][ids], m["epsilon"][ids]))
        r.fem_kernel()
        r.subit = r.subit + 1
        if not (r.subit < 20 and r.subres > r.subTol):
            break
    r.ibm_spread()
with open(out, "wb") as f:
    f.write(np.array([d["n_nodes"], d["n_bc"], d["n_ibm"], len(d["fm_node"]), len(calls)], "<i4").tobytes())
    f.write(np.array([d["alpha"], d["delta"], d["Dt"], d["Dm"], d["gravityX"], d["gravityY"], d["ref_L"]], "<f8").tobytes())
    for k in ("pos0", "angle0", "el"):
        f.write(np.ascontiguousarray(d[k], "<f8").tobytes())
    f.write(np.ascontiguousarray(d["pm_el"], "<i4").tobytes()); f.write(np.ascontiguousarray(d["pm_zeta"], "<f8").tobytes())
    f.write(np.ascontiguousarray(d["fm_first"], "<i4").tobytes()); f.write(np.ascontiguousarray(d["fm_node"], "<i4").tobytes())
    f.write(np.ascontiguousarray(d["fm_z1"], "<f8").tobytes()); f.write(np.ascontiguousarray(d["fm_z2"], "<f8").tobytes())
    for kind, t, relax, st, force, eps in calls:
        f.write(np.array([kind, t], "<i4").tobytes()); f.write(np.array([relax], "<f8").tobytes())
        f.write(np.ascontiguousarray(st, "<f8").tobytes()); f.write(np.ascontiguousarray(force, "<f8").tobytes())
        f.write(np.ascontiguousarray(eps, "<f8").tobytes())
r.close()
print("recorded %%d calls" %% len(calls))
'''


@pytest.fixture(scope="module")
def race_exe():
    os.makedirs(os.path.dirname(RACE_EXE), exist_ok=True)
    if not os.path.exists(RACE_EXE) or os.path.getmtime(RACE_EXE) < max(os.path.getmtime(RACE_SRC), os.path.getmtime(CORE)):
        r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fsanitize=thread", "-Wall", "-Wextra", "-o", RACE_EXE,
                            RACE_SRC, "-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("no ThreadSanitizer toolchain here: " + r.stderr[-300:])
    return RACE_EXE


@pytest.mark.parametrize("nthreads", [7, 32])
@pytest.mark.parametrize("case,steps,fb", [("InvertedFlag", 12, 0), ("PELskin", 5, 3)], ids=["InvertedFlag", "PELskin"])
def test_cta_of_real_threads_is_race_free_and_identical_to_serial(case, steps, fb, nthreads, race_exe, tmp_path):
    """Every recorded call of a live run, executed by one thread and by a CTA of `nthreads` real threads (FEM_SYNC = pthread barrier)
    under ThreadSanitizer: no data race reported, results bit-identical."""
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built (make -C oracle ref)" % case)
    calls = str(tmp_path / "calls.bin")
    p = subprocess.run([sys.executable, "-c", RECORD % dict(root=ROOT, case=case, steps=steps, fb=fb, out=calls)], capture_output=True,
                       text=True, timeout=900, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    q = subprocess.run([race_exe, calls, str(nthreads)], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, TSAN_OPTIONS="exitcode=66 halt_on_error=0"))
    assert "ThreadSanitizer" not in q.stderr, q.stderr[:4000]
    assert q.returncode == 0, q.stdout[-2000:] + q.stderr[-2000:]
    print(q.stdout.strip())
