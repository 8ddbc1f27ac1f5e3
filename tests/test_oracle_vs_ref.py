"""The oracle against the compiled, unmodified reference run side by side (oracle/_ref/libref_<case>.so).

Needs the reference objects built by `make -C oracle ref` (they travel to the GPU box; on a checkout without them these
tests skip and tests/test_oracle_golden.py carries the pin).  Each case runs in a subprocess: the reference keeps its
case in compile-time constants, deletes ./Results in its working directory and exit(99)s on its own errors.
"""
import os
import subprocess
import sys

import pytest

from oracle import refharness
from tests import cases as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
from oracle import oracle as O
from tests.cases import rel_l2
case, steps = %(case)r, %(steps)d
r = RefCase(case)
o = O.Oracle(O.params_from_ref(r))
assert (o.Dx, o.Dt, o.Dm, o.Drho) == (r.Dx, r.Dt, r.Dm, r.Drho)
for nm in ("f", "rho", "u", "force_xy", "u_in", "rho_in"):
    assert np.array_equal(o.get(nm), getattr(r, nm)()), "init " + nm
assert np.array_equal(o.types(), r.type())
assert np.array_equal(o.bcvec(), r.bcvec())
if r.n_markers:
    m = r.markers()
    o.set_markers(m["pos"], m["vel"], m["ds"], m["epsilon"])
    assert o.find_support() == 0
    for a, b in zip(o.supports(), r.supports()):
        assert np.array_equal(a, b)
exact = True          # both collision operators: the oracle is written in the reference's operation order (round 2: central moments too)
if not r.has_flex:
    r.step(steps); o.step(steps)
else:
    # flexible bodies: the reference's own FEM / epsilon code drives the markers, the oracle follows across the seam
    for s in range(steps):
        r.t = r.t + 1; r.lbm_kernel()
        o.t = r.t; o.lbm_kernel()
        r.subit = 0
        while True:
            r.recompute_object_vals()
            m = r.markers()
            o.set_markers(m["pos"], m["vel"], m["ds"], m["epsilon"]); o.find_support()
            r.ibm_interp(); o.ibm_interp()
            assert np.array_equal(o.marker_force(), r.markers()["force"])
            r.fem_kernel(); r.subit = r.subit + 1
            if not (r.subit < 20 and r.subres > r.subTol):
                break
        r.ibm_spread(); o.ibm_spread()
for nm in ("f", "rho", "u", "force_ibm"):
    a, b = o.get(nm), getattr(r, nm)()
    if exact:
        assert np.array_equal(a, b), nm
    else:
        assert rel_l2(a, b) < 1e-13, (nm, rel_l2(a, b))
print("OK")
'''

STEPS = {"Honami": 4, "InvertedFlag": 20, "PELskin": 20, "TurekHron": 40}


@pytest.mark.parametrize("case", K.ALL_CASES)
def test_side_by_side(case):
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built" % case)
    code = SCRIPT % dict(root=ROOT, case=case, steps=STEPS.get(case, 120))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_stream_map_bit_exact():
    """Push map recovered from the reference: with rho_n = 0 the BGK equilibrium is exactly zero, so one lbmKernel()
    pass over tagged populations gives f_new[recv, v] = tag - omega*tag; decoding the tags yields recv for every (node, v).
    The periodic case consumes every wrapped population, so the whole map incl. wrap-around is covered."""
    case = "t_periodic_bgk"
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built" % case)
    code = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
from oracle import oracle as O
r = RefCase(%(case)r)
o = O.Oracle(O.params_from_ref(r))
Nx, Ny = r.Nx, r.Ny
tags = np.arange(1, Nx * Ny * 9 + 1, dtype=np.float64).reshape(Nx, Ny, 9)
r.set_state(f=tags, rho=np.zeros((Nx, Ny)), u=np.zeros((Nx, Ny, 2)), force_xy=np.zeros((Nx, Ny, 2)))
r.t = 1; r.lbm_kernel()
out = r.f()
expect = tags + r.omega * (0.0 - tags)        # value the reference stores for a population carrying `tag`
lut = {float(expect[i, j, v]): (i * Ny + j, v) for i in range(Nx) for j in range(Ny) for v in range(9)}
assert len(lut) == Nx * Ny * 9
for i in range(Nx):
    for j in range(Ny):
        for v in range(9):
            src, sv = lut[float(out[i, j, v])]
            assert sv == v
            si, sj = divmod(src, Ny)
            assert o.stream_target(si, sj, v) == i * Ny + j
print("OK")
''' % dict(root=ROOT, case=case)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
