"""Pins the file-format restatements used as checkers of the device-fed file paths (include/life_b200.h, "device-fed files")
against the compiled, unmodified reference's OWN writers and reader (oracle/_ref/libref_<case>.so):

  oracle/fluidfiles.py  vti_bytes()      == GridClass::writeVTK      (src/Grid.cpp:790-898), byte for byte
  oracle/fluidfiles.py  restart_bytes()  == GridClass::writeRestart  (src/Grid.cpp:1163-1229), byte for byte
  life_vtk_frame()      (host half of life_write_vtk, no device needed) == the head / tail of the reference's .vti
  GridClass::readRestart accepts the restated bytes and recovers the state bit for bit

CPU only.  The GPU tests (tests/test_gpu_output.py) then hold life_write_vtk / life_write_restart / life_read_restart to these
same bytes.  Each case runs in a subprocess (the reference keeps one compile-time case per process).
"""
import os
import subprocess
import sys

import pytest

from oracle import refharness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
from tests import vtkfile as V, restartfile as R
from life_b200 import capi
case, steps = %(case)r, %(steps)d
r = RefCase(case)
for phase in ("initial", "stepped"):
    if phase == "stepped":
        r.step(steps)
    rho, u, f, fibm = r.rho(), r.u(), r.f(), r.force_ibm()
    # --- .vti ---
    ref = open(r.write_vtk(), "rb").read()
    mine = V.fluid_bytes(rho, u, r.Dx, r.Dt, r.Dm, r.Drho, r.rho_p, r.ref_P)
    assert len(ref) == len(mine), (phase, len(ref), len(mine))
    assert ref == mine, (phase, "vti bytes differ at", next(k for k in range(len(ref)) if ref[k] != mine[k]))
    head, tail = capi.vtk_frame(r.Nx, r.Ny, r.Dx)
    assert ref.startswith(head) and ref.endswith(tail)
    assert len(ref) == len(head) + 3 * 8 + 5 * 8 * r.Nx * r.Ny + len(tail)
    back = V.read_fluid(os.path.join(r.workdir, "Results", "VTK", "Fluid.%%d.vti" %% r.t), r.Nx, r.Ny)
    assert np.array_equal(back["density"], rho * r.Drho) and np.array_equal(back["velocity"][:, :, 0], u[:, :, 0] * (r.Dx / r.Dt))
    # --- Fluid.restart ---
    path = os.path.join(r.write_restart(), "Fluid.restart")
    ref = open(path, "rb").read()
    mine = R.fluid_bytes(r.t, r.omega, r.Dx, r.Dt, r.Dm, rho, u, fibm, f)
    assert ref == mine, (phase, "restart bytes differ")
# the reference's reader on bytes produced by the restatement, from a scrambled state
t_end = r.t
r.set_state(f=f * 0 + 7.0, rho=rho * 0 + 3.0, u=u * 0 - 1.0, force_ibm=fibm * 0 + 9.0)
open(path, "wb").write(mine)
assert r.read_restart() == t_end
for name, want in (("rho", rho), ("u", u), ("f", f), ("force_ibm", fibm)):
    assert np.array_equal(getattr(r, name)(), want), name
r.close()
print("OK")
'''

CASES = [("LidDrivenCavity", 20), ("ChannelFlow", 20), ("Cylinder", 5), ("t_womersley", 10), ("t_periodic_cm", 10)]


@pytest.mark.parametrize("case,steps", CASES, ids=[c for c, _ in CASES])
def test_format_restatements_match_the_reference_writers(case, steps):
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built (make -C oracle ref)" % case)
    p = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, case=case, steps=steps)], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-3000:] + p.stderr[-3000:]


def test_vtk_frame_arguments():
    from life_b200 import capi
    import ctypes as C
    L = capi.load()
    n = C.c_int64()
    assert L.life_vtk_frame(0, 5, 1.0, None, 0, C.byref(n), None, 0, None) == capi.E_ARG
    small = C.create_string_buffer(8)
    assert L.life_vtk_frame(5, 5, 1.0, small, 8, C.byref(n), None, 0, None) == capi.E_ARG     # buffer too small
    assert n.value > 8
    # offsets are 64-bit: 65536^2 nodes = 2^35 bytes per scalar block (the reference's int arithmetic overflows there)
    head, _ = capi.vtk_frame(65536, 65536, 1.0 / 65535)
    assert b'offset="%d"' % (65536 * 65536 * 8 + 8) in head and b'offset="%d"' % (2 * (65536 * 65536 * 8 + 8)) in head
