"""cfg.inplace: ONE population buffer (72 B/node resident), the sweep collides in place and streaming is an offset update
(csrc/lbm_bulk.cu: k_bulk_shift; ctx.h: PopShift).  Every consumer of the populations addresses them through the shifted layout,
so every path is held to the same bar as the two-buffer layout: the oracle / the compiled reference's fixtures, bitwise in exact
mode, and the device-fed files byte for byte."""
import os

import numpy as np
import pytest

from tests import cases as K

pytestmark = pytest.mark.gpu
LBM_CASES = K.EXAMPLES_LBM + K.EXTRA


def _steps(ctx, n, first=1):
    for t in range(first, first + n):
        ctx.step(t)


@pytest.mark.parametrize("case", LBM_CASES)
def test_inplace_fields_match_oracle_and_two_buffer_layout(case):
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    N = int(g["steps"]) | 1          # odd: the layout offsets are non-zero when the state is read back
    a = capi.Context(K.life_config(o.params, o, inplace=1))
    K.upload_from_oracle(a, o)
    _steps(a, N)
    b = capi.Context(K.life_config(o.params, o, kernel=1))
    K.upload_from_oracle(b, o)
    _steps(b, N)
    o.step(N)
    sa, sb = a.download_state(), b.download_state()
    for name in ("rho", "u", "f"):
        assert K.rel_l2(sa[name], o.get(name)) < K.TOL, (case, name)
        assert K.rel_l2(sa[name], sb[name], floor=1e-3 if name == "u" else 0.0) < 1e-13, (case, name)
    va, vb = a.max_speed(), b.max_speed()
    assert abs(va[0] - vb[0]) < 1e-13 and va[1] == vb[1]
    a.close()
    b.close()


@pytest.mark.parametrize("case", LBM_CASES)
def test_inplace_exact_mode_is_bitwise_the_reference(case):
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    N = int(g["steps"])
    a = capi.Context(K.life_config(o.params, o, inplace=1, exact=1))
    K.upload_from_oracle(a, o)
    _steps(a, N)
    o.step(N)
    st = a.download_state()
    for name in ("rho", "u", "f"):
        assert np.array_equal(st[name], o.get(name)), (case, name, K.rel_l2(st[name], o.get(name)))
        assert np.array_equal(K.sampled(st[name], g), g[name]), (case, "golden " + name)
    a.close()


@pytest.mark.parametrize("shape", [(40, 36), (37, 41), (5, 7), (130, 515)])
def test_inplace_push_map_bit_exact(shape):
    """omega = 0: a pure push.  After 1, 2 and 3 steps the tagged populations must sit where the reference's modulo sends them."""
    from life_b200 import capi
    from oracle import oracle as O
    Nx, Ny = shape
    p = O.Params(Nx=Nx, Ny=Ny, omega=1.0, wall_left=0, wall_right=0, wall_bottom=0, wall_top=0)
    o = O.Oracle(p)
    cfg = K.life_config(p, o, inplace=1)
    cfg.omega = 0.0
    ctx = capi.Context(cfg)
    tags = np.arange(1, Nx * Ny * 9 + 1, dtype=np.float64).reshape(Nx, Ny, 9)
    ctx.upload_state(tags)
    expect = tags
    for t in range(1, 4):
        ctx.step(t)
        nxt = np.zeros_like(tags)
        for v in range(9):
            tgt = np.array([[o.stream_target(i, j, v) for j in range(Ny)] for i in range(Nx)])
            nxt.reshape(-1, 9)[tgt.ravel(), v] = expect[:, :, v].ravel()
        expect = nxt
        assert np.array_equal(ctx.download_state()["f"], expect), t
    ctx.close()


@pytest.mark.parametrize("exact", [0, 1], ids=["fast", "exact"])
@pytest.mark.parametrize("case", K.EXAMPLES_IBM)
def test_inplace_fsi_trace_replay(case, exact):
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o, inplace=1, exact=exact))
    K.upload_from_oracle(ctx, o)
    steps = g["trace_step"]
    k = 0
    worst = 0.0
    for t in range(1, int(g["steps"]) + 1):
        ctx.step(t)
        while True:
            assert steps[k] == t
            ctx.ibm_set_markers(g["trace_pos"][k], g["trace_vel"][k], g["trace_ds"][k], g["trace_eps"][k])
            force = ctx.ibm_interp()
            if exact:
                assert np.array_equal(force, g["trace_force"][k]), (case, t, k)
            worst = max(worst, K.rel_l2(force, g["trace_force"][k], floor=1e-6))
            last = g["trace_last"][k]
            k += 1
            if last:
                break
        ctx.ibm_spread()
    assert worst < K.TOL, worst
    st = ctx.download_state()
    for name in ("rho", "u", "f", "force_ibm"):
        if exact:
            assert np.array_equal(K.sampled(st[name], g), g[name]), (case, name)
        else:
            assert K.rel_l2(K.sampled(st[name], g), g[name]) < K.TOL, (case, name)
    ctx.close()


@pytest.mark.parametrize("case", ["t_convective", "t_pressure_left", "Cylinder"])
def test_inplace_files_and_restart_roundtrip(case, tmp_path):
    """Device-fed Fluid.<t>.vti / Fluid.restart written from the shifted layout (sync and async snapshot) are the bytes the two-buffer
    layout writes in exact mode, and reading the restart file back continues the run identically."""
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    ctxs = []
    for inplace in (1, 0):
        c = capi.Context(K.life_config(o.params, o, inplace=inplace, exact=1))
        K.upload_from_oracle(c, o)
        _steps(c, 21)
        ctxs.append(c)
    a, b = ctxs
    files = {}
    for name, c in (("a", a), ("b", b)):
        for mode, tag in ((capi.IO_SYNC, "sync"), (capi.IO_ASYNC, "async")):
            if name == "b" and tag == "async":
                continue
            vti, rst = str(tmp_path / ("%s_%s.vti" % (name, tag))), str(tmp_path / ("%s_%s.restart" % (name, tag)))
            c.write_vtk(vti, o.params.rho_p, 0.25, mode)
            c.write_restart(rst, 21, mode)
            c.io_wait()
            files[(name, tag)] = (open(vti, "rb").read(), open(rst, "rb").read())
    assert files[("a", "sync")] == files[("b", "sync")]
    assert files[("a", "async")] == files[("b", "sync")]
    # restart from the file into a fresh in-place context: same continuation as the context that never stopped
    c2 = capi.Context(K.life_config(o.params, o, inplace=1, exact=1))
    t0 = c2.read_restart(str(tmp_path / "a_sync.restart"), o.get("force_xy").reshape(-1, 2)[0], o.get("u_in"), o.get("rho_in"))
    assert t0 == 21
    _steps(a, 10, first=22)
    _steps(c2, 10, first=22)
    sa, sc = a.download_state(), c2.download_state()
    for name in ("f", "rho", "u"):
        assert np.array_equal(sa[name], sc[name]), name
    for c in (a, b, c2):
        c.close()


def test_inplace_uses_one_population_buffer():
    """72 B/node resident: a 12288^2 lattice (87 GB with two buffers... 21.7 GB per buffer at 9 x 8 B) — check the device memory the context takes."""
    import torch
    from life_b200 import capi
    N = 8192
    free0 = torch.cuda.mem_get_info()[0]
    c = capi.Context(capi.Config(Nx=N, Ny=N, omega=1.0, wall_top=2, Dx=1.0, Dt=1.0, Dm=1.0, inplace=1))
    used1 = free0 - torch.cuda.mem_get_info()[0]
    c.close()
    c = capi.Context(capi.Config(Nx=N, Ny=N, omega=1.0, wall_top=2, Dx=1.0, Dt=1.0, Dm=1.0))
    used2 = free0 - torch.cuda.mem_get_info()[0]
    c.close()
    per_buffer = 9 * 8 * N * N
    assert used1 < 1.15 * per_buffer + (64 << 20), used1
    assert used2 > 1.9 * per_buffer, used2
