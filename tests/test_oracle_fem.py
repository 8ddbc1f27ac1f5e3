"""Pins oracle/life_oracle_fem.c — the C restatement of the reference's structural solver (groundwork for SURVEY.md §8f row 3,
DESIGN.md §10) — against the compiled, unmodified reference (oracle/_ref/libref_<case>.so) inside LIVE fluid-structure runs:
at every sub-iteration of every time step the reference's own loop (lbmKernel, recomputeObjectVals, ibmKernelInterp, femKernel,
ibmKernelSpread; src/Objects.cpp:26-60) advances the case, and next to it, for every flexible body,

  * resetValues + predictor (src/FEMBody.cpp:341-349, :259-289) / the Aitken-relaxed update (src/Objects.cpp:195-208), and
  * dynamicFEM (src/FEMBody.cpp:26-68: corotational beam elements, Newmark, Newton-Raphson over LAPACK)

are repeated by the restatement from the same state; marker positions and velocities, all eleven state vectors, the Newton-Raphson
iteration count and the sub-iteration residual sums must come out BIT FOR BIT the same.  CPU only.
"""
import os
import subprocess
import sys

import pytest

from oracle import refharness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
from oracle.oracle_fem import FemBody
case, steps = %(case)r, %(steps)d
r = RefCase(case)
nb = r.fem_count()
assert nb >= 1
desc = [r.fem_body(fb) for fb in range(nb)]
orc = [FemBody(d) for d in desc]
calls = nr_its = 0
def same(a, b, what):
    assert np.array_equal(np.asarray(a), np.asarray(b)), (what, float(np.abs(np.asarray(a) - np.asarray(b)).max()))
for step in range(steps):
    r.t = r.t + 1
    r.lbm_kernel()
    r.subit = 0
    while True:
        # ---- predictor (sub-iteration 0) or relaxed update, then supports / ds / epsilon: the reference ...
        before = [r.fem_get_state(fb, desc[fb]["n_dof"]) for fb in range(nb)]
        r.recompute_object_vals()
        m = r.markers()
        # ... and the restatement from the same state
        for fb in range(nb):
            orc[fb].set_state(before[fb])
            pos, vel = orc[fb].predict(r.t) if r.subit == 0 else orc[fb].relax(r.relax)
            ids = desc[fb]["marker"]
            same(pos, m["pos"][ids], "predict/relax pos"); same(vel, m["vel"][ids], "predict/relax vel")
            same(orc[fb].get_state(), r.fem_get_state(fb, desc[fb]["n_dof"]), "predict/relax state")
        r.ibm_interp()
        m = r.markers()
        # ---- dynamicFEM, body by body: restatement first, then the reference on a copy of the state that is put back afterwards
        #      (femKernel below is what really advances the run)
        for fb in range(nb):
            ids = desc[fb]["marker"]
            st = r.fem_get_state(fb, desc[fb]["n_dof"])
            orc[fb].set_state(st)
            pos, vel, res = orc[fb].dynamic(m["force"][ids], m["epsilon"][ids])
            ref_res = r.fem_dynamic(fb)
            after = r.fem_get_state(fb, desc[fb]["n_dof"])
            m2 = r.markers()
            assert res[4] == ref_res[4], ("Newton-Raphson iterations", res, ref_res)
            same(res[:4], ref_res[:4], "subRes, subNum, subDen, resNR")
            same(orc[fb].get_state()[:9], after[:9], "state after dynamicFEM")
            same(pos, m2["pos"][ids], "marker pos"); same(vel, m2["vel"][ids], "marker vel")
            r.fem_set_state(fb, st)
            calls += 1; nr_its += res[4]
        r.set_marker_posvel(m["pos"], m["vel"])
        r.fem_kernel()
        r.subit = r.subit + 1
        if not (r.subit < 20 and r.subres > r.subTol):
            break
    r.ibm_spread()
print("%%s: %%d bodies, %%d steps, %%d dynamicFEM calls (%%d Newton-Raphson iterations) bit-identical" %% (case, nb, steps, calls, nr_its))
r.close()
print("OK")
'''

CASES = [("TurekHron", 80), ("InvertedFlag", 25), ("PELskin", 12), ("Honami", 6)]


@pytest.mark.parametrize("case,steps", CASES, ids=[c for c, _ in CASES])
def test_fem_restatement_matches_the_reference_bit_for_bit(case, steps):
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built (make -C oracle ref)" % case)
    p = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, case=case, steps=steps)], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-3000:] + p.stderr[-3000:]
    print(p.stdout.strip().splitlines()[-2])
