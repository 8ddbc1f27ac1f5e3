"""First device run of the structural solver (life_fem_* of include/life_b200.h; csrc/fem.cu over csrc/fem_core.h; SURVEY.md §8f
row 3).  The solver's arithmetic and barrier placement are verified on the CPU (tests/test_fem_core.py) and it is compiled into
the library.  All four flexible examples have run on a B200 (round 1's round-end run, GPUTEST_r01.json) and are ordinary tests.
Each case runs in a subprocess so that a fault cannot poison the CUDA context of the other tests.

Method = tests/test_fem_core.py with the device in place of the serial host build: inside live fluid-structure runs of the
compiled reference, every predictor / relaxed update / dynamicFEM call of every flexible body is repeated through the C ABI from
the reference's state; marker positions and velocities (read back with life_ibm_get_markers), state vectors and residual sums must
agree to rounding (own LU instead of LAPACK: 1e-11 of the body length for displacements, 1e-8 for velocities / accelerations /
residual sums).
"""
import os
import subprocess
import sys

import pytest

from oracle import refharness

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
from life_b200 import capi
from tests import cases as K
case, steps = %(case)r, %(steps)d
r = RefCase(case)
g = K.golden(case)
o = K.make_oracle(g)
ctx = capi.Context(K.life_config(o.params, o, device=0))
assert (ctx.cfg.Dt, ctx.cfg.Dm) == (r.Dt, r.Dm)
nb = r.fem_count()
desc = [r.fem_body(fb) for fb in range(nb)]
def send(m):
    ctx.ibm_set_markers(m["pos"], m["vel"], m["ds"], m["epsilon"])
    ctx.ibm_set_forces(m["force"])
send(r.markers())
ctx.fem_create(desc)
worst = {}
def close(a, b, what, scale_floor, tol=1e-11):
    a, b = np.asarray(a, float), np.asarray(b, float)
    err = float(np.abs(a - b).max() / max(np.abs(b).max(), scale_floor))
    worst[what] = max(worst.get(what, 0.0), err)
    assert err < tol, (what, err)
calls = 0
for step in range(steps):
    r.t = r.t + 1
    r.lbm_kernel()
    r.subit = 0
    while True:
        # ---- predictor / relaxed update of every body in one launch
        m0 = r.markers()
        for fb in range(nb):
            ctx.fem_set_state(fb, r.fem_get_state(fb, desc[fb]["n_dof"]))
        send(m0)
        r.recompute_object_vals()
        m = r.markers()
        if r.subit == 0:
            ctx.fem_predict(r.t)
        else:
            ctx.fem_relax(r.relax)
        pos, vel = ctx.ibm_get_markers()
        rigid = np.ones(len(pos), bool)
        for fb in range(nb):
            d, ids = desc[fb], desc[fb]["marker"]
            rigid[ids] = False
            close(pos[ids], m["pos"][ids], "predict/relax marker pos", d["ref_L"])
            close(vel[ids], m["vel"][ids], "predict/relax marker vel", d["ref_L"] / d["Dt"] * 1e-3, 1e-8)
            got, ref_st = ctx.fem_get_state(fb), r.fem_get_state(fb, d["n_dof"])
            for k in (0, 3, 6, 9, 10):
                close(got[k], ref_st[k], "predict/relax displacement vectors", d["ref_L"])
            for k, s in ((1, 1.0 / d["Dt"]), (2, 1.0 / d["Dt"] ** 2), (4, 1.0 / d["Dt"]), (5, 1.0 / d["Dt"] ** 2)):
                close(got[k], ref_st[k], "predict/relax velocity / acceleration vectors", d["ref_L"] * s * 1e-3, 1e-8)
        assert np.array_equal(pos[rigid], m0["pos"][rigid]) and np.array_equal(vel[rigid], m0["vel"][rigid])   # other markers untouched
        # ---- dynamicFEM of every body in one launch
        r.ibm_interp()
        m = r.markers()
        before = [r.fem_get_state(fb, desc[fb]["n_dof"]) for fb in range(nb)]
        for fb in range(nb):
            ctx.fem_set_state(fb, before[fb])
        send(m)
        sums, per = ctx.fem_dynamic()
        pos, vel = ctx.ibm_get_markers()
        ref_sums = np.zeros(3)
        for fb in range(nb):
            d, ids = desc[fb], desc[fb]["marker"]
            ref_res = r.fem_dynamic(fb)
            after = r.fem_get_state(fb, d["n_dof"])
            m2 = r.markers()
            ref_sums += np.array(ref_res[:3])
            assert abs(int(per[fb, 4]) - ref_res[4]) <= 1, ("Newton-Raphson iterations", per[fb], ref_res)
            got = ctx.fem_get_state(fb)
            close(got[0], after[0], "U after dynamicFEM", d["ref_L"])
            close(got[1], after[1], "Udot after dynamicFEM", d["ref_L"] / d["Dt"] * 1e-3, 1e-8)
            close(got[2], after[2], "Udotdot after dynamicFEM", d["ref_L"] / d["Dt"] ** 2 * 1e-3, 1e-8)
            close(got[7], after[7], "R_k", d["ref_L"]); close(got[8], after[8], "R_km1", d["ref_L"])
            close(pos[ids], m2["pos"][ids], "marker pos", d["ref_L"])
            close(vel[ids], m2["vel"][ids], "marker vel", d["ref_L"] / d["Dt"] * 1e-3, 1e-8)
            close(per[fb, :3], ref_res[:3], "subRes, subNum, subDen", d["ref_L"] ** 2 * 1e-6, 1e-8)
            r.fem_set_state(fb, before[fb])
            calls += 1
        close(sums, ref_sums, "sums over the bodies", desc[0]["ref_L"] ** 2 * 1e-6, 1e-8)
        r.set_marker_posvel(m["pos"], m["vel"])
        r.fem_kernel()
        r.subit = r.subit + 1
        if not (r.subit < 20 and r.subres > r.subTol):
            break
    r.ibm_spread()
print("%%s: %%d bodies, %%d steps, %%d dynamicFEM calls on the device, %%d kernel launches; worst relative differences: %%s"
      %% (case, nb, steps, calls, ctx.launch_count(), {k: float("%%.1e" %% v) for k, v in worst.items()}))
ctx.close(); r.close()
print("OK")
'''

CASES = [pytest.param("InvertedFlag", 25, id="InvertedFlag"), pytest.param("Honami", 6, id="Honami"),
         pytest.param("TurekHron", 40, id="TurekHron"), pytest.param("PELskin", 12, id="PELskin")]


@pytest.mark.parametrize("case,steps", CASES)
def test_device_structural_solver_matches_the_reference(case, steps):
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built" % case)
    p = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, case=case, steps=steps)], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    print(p.stdout[-1500:])
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-3000:] + p.stderr[-3000:]


RESIDENT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
from life_b200 import capi
from tests import cases as K
case, steps = %(case)r, %(steps)d
r = RefCase(case)
g = K.golden(case)
o = K.make_oracle(g)
ctx = capi.Context(K.life_config(o.params, o, device=0))
nb = r.fem_count()
desc = [r.fem_body(fb) for fb in range(nb)]
m = r.markers()
n = len(m["ds"])
ctx.ibm_set_markers(m["pos"], m["vel"], m["ds"], m["epsilon"])
ctx.ibm_set_forces(m["force"])
ctx.fem_create(desc)
uni = bool(r.flags & RefCase.FLAG_UNI_EPS)
groups = [np.arange(n)] if uni else [np.asarray(d["marker"]) for d in desc]
worst = {}
def close(a, b, what, floor, tol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    err = float(np.abs(a - b).max() / max(np.abs(b).max(), floor))
    worst[what] = max(worst.get(what, 0.0), err)
    assert err < tol, (what, err)
subits = 0
for step in range(steps):
    r.t = r.t + 1
    r.lbm_kernel()
    # the GPU lattice takes the reference's mid-step state (post-stream populations, IBM force not yet spread)
    ctx.upload_state(r.f(), None, None, r.force_xy(), None, r.u_in(), r.rho_in())
    r.subit = 0
    while True:
        # both sides start the sub-iteration from the reference's state
        for fb in range(nb):
            ctx.fem_set_state(fb, r.fem_get_state(fb, desc[fb]["n_dof"]))
        m0 = r.markers()
        ctx.ibm_set_markers(m0["pos"], m0["vel"], m0["ds"], m0["epsilon"])
        ctx.ibm_set_forces(m0["force"])
        r.recompute_object_vals()                      # predictor / relax + findSupport + computeDs + computeEpsilon on the host
        ctx.fsi_move(r.t, r.subit, r.relax)            # the same on the device, from the device's own marker arrays
        ctx.ibm_compute_epsilon(groups)
        m1 = r.markers()
        pos, vel = ctx.ibm_get_markers()
        force, ds, eps = ctx.ibm_get_marker_state()
        L = desc[0]["ref_L"]
        close(pos, m1["pos"], "marker pos after move", L, 1e-11)
        close(ds, m1["ds"], "ds", 1.0, 1e-9)
        close(eps, m1["epsilon"], "epsilon (device LU vs LAPACK)", 1.0, 1e-7)
        cnt, idx, jdx, dirac = ctx.ibm_get_supports()
        rc, ri, rj, rd = r.supports()
        same = np.array_equal(cnt, rc) and np.array_equal(idx, ri) and np.array_equal(jdx, rj)
        worst["markers with a different support set"] = worst.get("markers with a different support set", 0) + (0 if same else int((cnt != rc).sum() + (idx != ri).any(axis=1).sum()))
        if same:
            close(dirac, rd, "delta weights", 1.0, 1e-9)
        r.ibm_interp()
        sums, per = ctx.fsi_force()                    # interp + dynamicFEM, forces never leave the device
        f_dev = ctx.ibm_get_marker_state()[0]
        close(f_dev, r.markers()["force"], "marker force (interp)", 1e-6, 1e-7)
        r.fem_kernel()
        ref_sub = r.subres
        sub = np.sqrt(sums[0]) / (L * np.sqrt(float(desc[0]["sim_dofs"])))
        close([sub], [ref_sub], "subRes", 1e-9, 1e-5)
        pos2, _ = ctx.ibm_get_markers()
        close(pos2, r.markers()["pos"], "marker pos after dynamicFEM", L, 1e-10)
        subits += 1
        r.subit = r.subit + 1
        if not (r.subit < 20 and r.subres > r.subTol):
            break
    r.ibm_spread()
assert worst.get("markers with a different support set", 0) == 0, worst
print("%%s: %%d steps, %%d sub-iterations resident on the device; worst relative differences: %%s" %% (case, steps, subits, {k: float("%%.1e" %% v) for k, v in worst.items()}))
ctx.close(); r.close()
print("OK")
'''


@pytest.mark.parametrize("case,steps", [("InvertedFlag", 8), ("Honami", 3), ("TurekHron", 10), ("PELskin", 4)])
def test_resident_subiteration_loop_matches_the_reference(case, steps):
    """life_fsi_move / life_ibm_compute_epsilon / life_fsi_force — the sub-iteration loop with the markers resident on the device — inside
    live FSI runs of the compiled reference: after every device call the marker positions, ds, epsilon, supports, forces and the
    residual agree with what the reference's host code (recomputeObjectVals, ibmKernelInterp, femKernel) produced from the same state."""
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built" % case)
    p = subprocess.run([sys.executable, "-c", RESIDENT % dict(root=ROOT, case=case, steps=steps)], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    print(p.stdout[-1500:])
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-3000:] + p.stderr[-3000:]
