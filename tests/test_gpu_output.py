"""The device-fed file paths (include/life_b200.h "device-fed files") against the reference's formats.

life_write_vtk / life_write_restart must produce, BYTE FOR BYTE, what GridClass::writeVTK (src/Grid.cpp:790-898) and
GridClass::writeRestart (src/Grid.cpp:1163-1229) write from the same rho / u / f / force_ibm.  The expected bytes come from
oracle/fluidfiles.py (the numpy restatement of both formats), which tests/test_output_files.py pins against the compiled reference's own writers
on the CPU; one test here also runs the compiled reference's writer and reader directly on the device's state / files.
life_read_restart must be the exact inverse of life_write_restart and fail with the reference's messages.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import cases as K
from tests import restartfile as R
from tests import vtkfile as V

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = ["LidDrivenCavity", "ChannelFlow", "Cylinder", "t_womersley", "t_periodic_cm", "t_convective"]


def _start(case, **cfg_kw):
    """context + oracle in the case's initial state (uploaded with rho / u, as a LIFE host does after initialiseGrid)"""
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o, **cfg_kw))
    K.upload_from_oracle(ctx, o)
    if case in K.EXAMPLES_IBM:
        ctx.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    return g, o, ctx


def _advance(case, ctx, t0, n):
    for t in range(t0 + 1, t0 + n + 1):
        ctx.step(t)
        if case in K.EXAMPLES_IBM:      # rigid body: markers fixed; leaves a non-zero force_ibm field behind
            ctx.ibm_interp()
            ctx.ibm_spread()
    return t0 + n


def _expected(o, st, t):
    p = o.params
    vti = V.fluid_bytes(st["rho"], st["u"], o.Dx, o.Dt, o.Dm, o.Drho, p.rho_p, 0.25)
    rst = R.fluid_bytes(t, p.omega, o.Dx, o.Dt, o.Dm, st["rho"], st["u"], st["force_ibm"], st["f"])
    return vti, rst


def _first_diff(a, b):
    if len(a) != len(b):
        return "lengths %d / %d" % (len(a), len(b))
    x, y = np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8)
    k = int(np.argmax(x != y))
    return "first differing byte %d of %d" % (k, len(a))


@pytest.mark.parametrize("staging", [None, 4096], ids=["staging-32MiB", "staging-min"])
@pytest.mark.parametrize("mode", [0, 1], ids=["sync", "async"])
@pytest.mark.parametrize("case", CASES)
def test_files_are_byte_identical_to_the_reference_format(case, mode, staging, tmp_path):
    g, o, ctx = _start(case)
    if staging:
        ctx.io_set_staging(staging)       # forces one .vti row / one restart column per chunk: many chunks, both slots reused
    vti_path, rst_path = str(tmp_path / "Fluid.7.vti"), str(tmp_path / "Fluid.restart")
    t = 0
    for phase in ("initial (uploaded rho, u)", "after steps"):
        if phase == "after steps":
            t = _advance(case, ctx, t, 12)
        st = ctx.download_state()
        want_vti, want_rst = _expected(o, st, t)
        l0 = ctx.launch_count()
        ctx.write_vtk(vti_path, o.params.rho_p, 0.25, mode)
        ctx.io_wait()
        assert ctx.launch_count() - l0 >= 3                              # at least one pack kernel per block, counted with the rest
        got = open(vti_path, "rb").read()
        assert got == want_vti, (case, phase, ".vti", _first_diff(got, want_vti))
        sec, nbytes, was_async = ctx.io_stats()
        assert nbytes == len(want_vti) and was_async == bool(mode)
        ctx.write_restart(rst_path, t, mode)
        ctx.io_wait()
        assert not os.path.exists(rst_path + ".temp")                  # renamed into place, src/Grid.cpp:1228
        got = open(rst_path, "rb").read()
        assert got == want_rst, (case, phase, "Fluid.restart", _first_diff(got, want_rst))
        assert ctx.io_stats()[1] == len(want_rst)
    if case == "Cylinder":
        assert np.abs(st["force_ibm"]).max() > 0          # the records really carried an IBM force
    ctx.close()


@pytest.mark.parametrize("case", ["ChannelFlow", "Cylinder"])
def test_async_write_freezes_the_state_while_the_steps_go_on(case, tmp_path):
    g, o, ctx = _start(case)
    ctx.io_set_staging(4096)
    t = _advance(case, ctx, 0, 10)
    st = ctx.download_state()
    want_vti, want_rst = _expected(o, st, t)
    ctx.write_restart(str(tmp_path / "Fluid.restart"), t, 1)
    t = _advance(case, ctx, t, 25)                       # the state moves on while the worker is still writing
    ctx.write_vtk(str(tmp_path / "a.vti"), o.params.rho_p, 0.25, 1)     # implies the wait for the restart job
    assert open(tmp_path / "Fluid.restart", "rb").read() == want_rst
    st2 = ctx.download_state()
    t = _advance(case, ctx, t, 5)
    ctx.io_wait()
    assert ctx.io_stats()[2] is True and not ctx.io_busy()
    want_vti2, _ = _expected(o, st2, t)
    assert open(tmp_path / "a.vti", "rb").read() == want_vti2
    assert not np.array_equal(st["f"], st2["f"])
    # and the stepping itself was not disturbed by the concurrent snapshot reads
    o2 = K.make_oracle(g)
    if case not in K.EXAMPLES_IBM:
        o2.step(t)
        now = ctx.download_state()
        for name in ("rho", "u", "f"):
            assert K.rel_l2(now[name], o2.get(name)) < K.TOL
    ctx.close()


@pytest.mark.parametrize("staging", [None, 4096], ids=["staging-32MiB", "staging-min"])
@pytest.mark.parametrize("case", ["LidDrivenCavity", "ChannelFlow", "Cylinder", "t_womersley", "t_convective"])
def test_read_restart_is_the_inverse_and_the_run_continues_bit_for_bit(case, staging, tmp_path):
    from life_b200 import capi
    g, o, a = _start(case)
    t = _advance(case, a, 0, 15)
    path = str(tmp_path / "Fluid.restart")
    a.write_restart(path, t)
    sa = a.download_state()
    fxy = o.get("force_xy").reshape(-1, 2)[0]
    b = capi.Context(K.life_config(o.params, o))
    if staging:
        b.io_set_staging(staging)
    assert b.read_restart(path, fxy, o.get("u_in"), o.get("rho_in")) == t
    sb = b.download_state()
    for name in ("f", "rho", "u", "force_ibm"):
        assert np.array_equal(sa[name], sb[name]), (case, name)
    if case in K.EXAMPLES_IBM:
        b.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    # continuing from the file is continuing the original run
    _advance(case, a, t, 10)
    _advance(case, b, t, 10)
    sa, sb = a.download_state(), b.download_state()
    for name in ("f", "rho", "u", "force_ibm"):
        if case == "t_womersley":
            # the first step after a restart takes u_n / rho_n from the file, an uninterrupted run evaluates them from f:
            # same numbers up to rounding, and force_xy of the step before is re-derived
            assert K.rel_l2(sb[name], sa[name], floor=1e-12) < K.TOL, (case, name)
        else:
            assert K.rel_l2(sb[name], sa[name], floor=1e-12) < 1e-12, (case, name)
    a.close()
    b.close()


def test_read_restart_fails_like_the_reference(tmp_path):
    from life_b200 import capi
    g, o, a = _start("ChannelFlow")
    t = _advance("ChannelFlow", a, 0, 3)
    path = str(tmp_path / "Fluid.restart")
    a.write_restart(path, t)
    raw = bytearray(open(path, "rb").read())
    a.close()

    def attempt(cfg_kw=None, data=None, name=path):
        if data is not None:
            open(name, "wb").write(data)
        ctx = capi.Context(K.life_config(o.params, o, **(cfg_kw or {})))
        try:
            with pytest.raises(capi.LifeError) as e:
                ctx.read_restart(name)
            with pytest.raises(capi.LifeError):      # and no half-read state is left usable
                ctx.step(1)
            return e.value
        finally:
            ctx.close()

    e = attempt(cfg_kw=dict(omega=o.params.omega * 1.0000001))                      # src/Grid.cpp:1103-1104
    assert e.code == capi.E_ARG and "Grid size/scaling has changed between runs...this is not supported" in str(e)
    bad = bytearray(raw)
    rec = 44 + 120 * (3 * o.Ny + 5)
    bad[rec + 4:rec + 8] = np.int32(6).tobytes()                                     # record (3, 5) claims j = 6
    e = attempt(data=bad, name=str(tmp_path / "bad_index"))                          # src/Grid.cpp:1137-1138
    assert e.code == capi.E_ARG and "Grid indices do not match Fluid.restart file...exiting" in str(e)
    e = attempt(data=raw[:len(raw) // 2], name=str(tmp_path / "short"))
    assert e.code == capi.E_IO
    e = attempt(name=str(tmp_path / "does_not_exist"))                               # src/Grid.cpp:1079-1080
    assert e.code == capi.E_IO and "Error opening Fluid.restart file...exiting" in str(e)


def test_write_errors_are_reported(tmp_path):
    from life_b200 import capi
    g, o, ctx = _start("LidDrivenCavity")
    with pytest.raises(capi.LifeError) as e:
        ctx.write_vtk(str(tmp_path / "no_such_dir" / "x.vti"), 1.0)
    assert e.value.code == capi.E_IO
    ctx.write_restart(str(tmp_path / "no_such_dir" / "Fluid.restart"), 1, 1)         # asynchronous: reported by the wait
    with pytest.raises(capi.LifeError) as e:
        ctx.io_wait()
    assert e.value.code == capi.E_IO
    ctx.write_vtk(str(tmp_path / "ok.vti"), 1.0)                                      # the context stays usable
    fresh = capi.Context(K.life_config(o.params, o))
    with pytest.raises(capi.LifeError) as e:
        fresh.write_vtk(str(tmp_path / "y.vti"), 1.0)                                 # nothing uploaded yet
    assert e.value.code == capi.E_STATE
    fresh.close()
    ctx.close()


REF_SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle.refharness import RefCase
from life_b200 import capi
from tests import cases as K
case = %(case)r
r = RefCase(case)
g = K.golden(case)
o = K.make_oracle(g)
ctx = capi.Context(K.life_config(o.params, o))
ctx.upload_state(r.f(), r.rho(), r.u(), r.force_xy(), r.force_ibm(), r.u_in(), r.rho_in())
if r.n_markers:
    m = r.markers()
    ctx.ibm_set_markers(m["pos"], m["vel"], m["ds"], m["epsilon"])
for t in range(1, 21):
    ctx.step(t)
    if r.n_markers:
        ctx.ibm_interp(); ctx.ibm_spread()
st = ctx.download_state()
# 1. the reference's own writers on the device's state == the device-fed writers
r.t = 20
r.set_state(f=st["f"], rho=st["rho"], u=st["u"], force_ibm=st["force_ibm"])
ref_vti = open(r.write_vtk(), "rb").read()
ref_rst = open(os.path.join(r.write_restart(), "Fluid.restart"), "rb").read()
mine = os.path.join(r.workdir, "mine")
os.makedirs(mine)
for mode in (0, 1):
    ctx.write_vtk(os.path.join(mine, "Fluid.20.vti"), r.rho_p, r.ref_P, mode)
    ctx.write_restart(os.path.join(mine, "Fluid.restart"), 20, mode)
    ctx.io_wait()
    assert open(os.path.join(mine, "Fluid.20.vti"), "rb").read() == ref_vti, "vti"
    assert open(os.path.join(mine, "Fluid.restart"), "rb").read() == ref_rst, "restart"
# 2. the reference's own reader accepts the device-written file
r.set_state(f=st["f"] * 0, rho=st["rho"] * 0, u=st["u"] * 0, force_ibm=st["force_ibm"] * 0 + 1)
os.replace(os.path.join(mine, "Fluid.restart"), os.path.join(r.workdir, "Results", "Restart", "Fluid.restart"))
assert r.read_restart() == 20
for name in ("f", "rho", "u", "force_ibm"):
    assert np.array_equal(getattr(r, name)(), st[name]), name
# 3. the device reader accepts the reference-written file
open(os.path.join(mine, "ref.restart"), "wb").write(ref_rst)
b = capi.Context(K.life_config(o.params, o))
assert b.read_restart(os.path.join(mine, "ref.restart"), r.force_xy().reshape(-1, 2)[0], r.u_in(), r.rho_in()) == 20
sb = b.download_state()
for name in ("f", "rho", "u", "force_ibm"):
    assert np.array_equal(sb[name], st[name]), name
b.close(); ctx.close(); r.close()
print("OK")
'''


@pytest.mark.parametrize("case", ["LidDrivenCavity", "Cylinder", "t_womersley"])
def test_against_the_compiled_reference_writer_and_reader(case):
    from oracle import refharness
    if not refharness.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built" % case)
    p = subprocess.run([sys.executable, "-c", REF_SCRIPT % dict(root=ROOT, case=case)], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("mode", [0, 1], ids=["sync", "async"])
def test_mid_size_lattice_default_staging(mode, tmp_path):
    """2048 x 1536 (3.1 M nodes; .vti 126 MB in 7 chunks, restart 377 MB in 12): ragged tiles in both directions, both staging
    slots in use, the worker overlapping the time loop."""
    from life_b200 import capi
    from oracle import oracle as O
    from tests.initstate import wavy_state
    Nx, Ny = 2048 + 17, 1536 - 5
    p = O.Params(Nx=Nx, Ny=Ny, omega=1.3, wall_left=2, wall_right=4, wall_bottom=1, wall_top=3, uxInlet_p=1.0, height_p=1.0,
                 rho_p=1.0, nu_p=0.01)
    o = O.Oracle(p)
    f0, rho0, u0 = wavy_state(Nx, Ny, False, amp=0.02, non_equilibrium=0.01)
    ctx = capi.Context(K.life_config(p, o))
    ctx.upload_state(f0, rho0, u0, None, None, o.get("u_in"), o.get("rho_in"))
    ctx.step_n(1, 20)
    st = ctx.download_state()
    want_vti, want_rst = _expected(o, st, 20)
    ctx.write_vtk(str(tmp_path / "a.vti"), p.rho_p, 0.25, mode)
    if mode:
        ctx.step_n(21, 50)
    ctx.io_wait()
    vti_stats = ctx.io_stats()
    ctx.write_restart(str(tmp_path / "Fluid.restart"), 20, mode)
    if mode:
        ctx.step_n(71, 50)
    ctx.io_wait()
    rst_stats = ctx.io_stats()
    print("\n%dx%d %s: .vti %.0f MB in %.3f s, restart %.0f MB in %.3f s" %
          (Nx, Ny, "async" if mode else "sync", vti_stats[1] / 1e6, vti_stats[0], rst_stats[1] / 1e6, rst_stats[0]))
    assert open(tmp_path / "a.vti", "rb").read() == want_vti
    if not mode:
        assert open(tmp_path / "Fluid.restart", "rb").read() == want_rst
    else:       # the restart snapshot was taken 50 steps later
        back = R.read_fluid(str(tmp_path / "Fluid.restart"))
        assert not np.array_equal(back["f"], st["f"])
        b = capi.Context(K.life_config(p, o))
        b.upload_state(f0, rho0, u0, None, None, o.get("u_in"), o.get("rho_in"))
        b.step_n(1, 70)
        sb = b.download_state()
        for name in ("f", "rho", "u"):
            assert np.array_equal(back[name], sb[name]), name
        b.close()
    ctx.close()


def test_markers_sent_before_the_next_step_do_not_drop_the_spread_force():
    """The reference keeps force_ibm from ibmKernelSpread until the next ibmKernelInterp zeroes it (src/Objects.cpp:105): sending
    the markers again in between (what the host does when it recomputes epsilon right after a restart, src/Objects.cpp:1184-1185)
    must not clear the force the next lbmKernel is about to use."""
    g, o, a = _start("Cylinder")
    _, _, b = _start("Cylinder")
    for t in range(1, 9):
        for ctx in (a, b):
            ctx.step(t)
            ctx.ibm_interp()
            ctx.ibm_spread()
        b.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])       # between spread and the next step
    sa, sb = a.download_state(), b.download_state()
    assert np.abs(sa["force_ibm"]).max() > 0
    for name in ("f", "rho", "u", "force_ibm"):
        assert np.array_equal(sa[name], sb[name]), name
    a.close()
    b.close()
