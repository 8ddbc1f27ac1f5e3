"""cfg.exact: the lattice-Boltzmann step in the reference's operation order (kernels of namespace life::exact, compiled with
-fmad=false) must reproduce the reference's doubles BIT FOR BIT — the reference's own regression protocol is `diff -r`
(testing/run-tests.sh:100).  Checked here through the C ABI against
  * the oracle (oracle/life_oracle.c, whose BGK path is itself bit-identical to the compiled reference, tests/test_oracle_vs_ref.py),
  * the committed fixtures the compiled reference wrote (tests/golden/*.npz): sampled fields, and for the body cases every marker
    force of every recorded sub-iteration,
  * the compiled reference itself (oracle/_ref/libref_<case>.so), stepped side by side in this process.
Both collision operators: BGK as collide_bgk_ref, central moments as collide_cm_ref (the reference's nine expanded polynomials,
src/Grid.cpp:143-223, term by term) — the oracle restates both in the reference's order as well and is bit-identical to the compiled
reference for both (tests/test_oracle_vs_ref.py, on the CPU).
"""
import numpy as np
import pytest

from tests import cases as K

pytestmark = pytest.mark.gpu

LBM_CASES = K.EXAMPLES_LBM + K.EXTRA
BGK_CASES = [c for c in LBM_CASES if not int(K.golden(c)["central_moments"])]
CM_CASES = [c for c in LBM_CASES if int(K.golden(c)["central_moments"])]


def _run(g, steps, **cfg_kw):
    from life_b200 import capi
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o, exact=1, **cfg_kw))
    K.upload_from_oracle(ctx, o)
    for t in range(1, steps + 1):
        ctx.step(t)
    o.step(steps)
    st = ctx.download_state()
    ctx.close()
    return o, st


@pytest.mark.parametrize("kernel", [1, 2], ids=["direct", "shuffle"])
@pytest.mark.parametrize("case", BGK_CASES)
def test_exact_bgk_step_is_bitwise_the_reference(case, kernel):
    g = K.golden(case)
    o, st = _run(g, int(g["steps"]), kernel=kernel)
    for name in ("rho", "u", "f"):
        assert np.array_equal(st[name], o.get(name)), (case, name, K.rel_l2(st[name], o.get(name)))
        # the compiled reference's own numbers (fixture written by oracle/_ref, tests/golden/make_golden.py)
        assert np.array_equal(K.sampled(st[name], g), g[name]), (case, "golden " + name)


@pytest.mark.parametrize("case", CM_CASES)
def test_exact_cm_step_is_bitwise_the_reference(case):
    g = K.golden(case)
    o, st = _run(g, int(g["steps"]))
    _, st2 = _run(g, int(g["steps"]), kernel=1)
    for name in ("rho", "u", "f"):
        assert np.array_equal(st[name], o.get(name)), (case, name)                          # the oracle (reference-ordered since round 2)
        assert np.array_equal(K.sampled(st[name], g), g[name]), (case, "golden " + name)   # the compiled reference's fixture: bitwise
        assert np.array_equal(st[name], st2[name]), (case, name)      # direct and shuffle kernels: same bits


@pytest.mark.parametrize("case", K.EXAMPLES_IBM)
def test_exact_fsi_trace_replay_is_bitwise(case):
    """LBM step (exact) on the GPU, then every recorded sub-iteration of the compiled reference's live FSI run: the host state the
    reference's FEM / epsilon code handed to ibmKernelInterp goes in, the marker forces that come back must be the reference's
    doubles; after the last step the sampled fields must be too.  All five body examples are BGK + ORDERED."""
    from life_b200 import capi
    g = K.golden(case)
    assert not int(g["central_moments"]) and int(g["ordered"])
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o, exact=1))
    K.upload_from_oracle(ctx, o)
    steps = g["trace_step"]
    k = 0
    for t in range(1, int(g["steps"]) + 1):
        ctx.step(t)
        while True:
            assert steps[k] == t
            ctx.ibm_set_markers(g["trace_pos"][k], g["trace_vel"][k], g["trace_ds"][k], g["trace_eps"][k])
            force = ctx.ibm_interp()
            assert np.array_equal(force, g["trace_force"][k]), (case, t, k, K.rel_l2(force, g["trace_force"][k], floor=1e-6))
            last = g["trace_last"][k]
            k += 1
            if last:
                break
        ctx.ibm_spread()
    st = ctx.download_state()
    for name in ("rho", "u", "f", "force_ibm"):
        assert np.array_equal(K.sampled(st[name], g), g[name]), (case, name)
    ctx.close()


@pytest.mark.parametrize("case", ["ChannelFlow", "t_convective", "t_womersley", "Cylinder", "LidDrivenCavity", "t_periodic_cm", "t_freeslip_cm"])
def test_exact_step_side_by_side_with_the_compiled_reference(case):
    """The unmodified reference (oracle/_ref/libref_<case>.so) and the GPU advance the same state step by step; whole fields are
    compared bit for bit after every step (Cylinder: with its rigid body, interp + spread included)."""
    from life_b200 import capi
    from oracle import refharness as RH
    if not RH.available(case):
        pytest.skip("oracle/_ref/libref_%s.so not built" % case)
    g = K.golden(case)
    o = K.make_oracle(g)          # only for parameters and the initial state
    ref = RH.RefCase(case)
    try:
        if int(g["wavy"]):
            ref.set_state(f=o.get("f"), rho=o.get("rho"), u=o.get("u"))
        ctx = capi.Context(K.life_config(o.params, o, exact=1))
        ctx.upload_state(ref.f(), ref.rho(), ref.u(), ref.force_xy(), ref.force_ibm(), ref.u_in(), ref.rho_in())
        body = ref.has_ibm
        if body:
            m = ref.markers()
            ctx.ibm_set_markers(m["pos"], m["vel"], m["ds"], m["epsilon"])
        for t in range(1, 26):
            ref.step(1)
            ctx.step(t)
            if body:
                force = ctx.ibm_interp()
                ctx.ibm_spread()
                assert np.array_equal(force, ref.markers()["force"]), (case, t)
            if t % 5 == 0 or t < 3:
                st = ctx.download_state()
                assert np.array_equal(st["f"], ref.f()), (case, t, "f")
                assert np.array_equal(st["rho"], ref.rho()), (case, t, "rho")
                assert np.array_equal(st["u"], ref.u()), (case, t, "u")
                assert np.array_equal(st["force_ibm"], ref.force_ibm()), (case, t, "force_ibm")
        ctx.close()
    finally:
        ref.close()
