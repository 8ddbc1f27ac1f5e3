"""Parity of the CUDA immersed-boundary kernels (support search, interpolate + force, spread) with the oracle and with
the call-site traces the compiled reference recorded (tests/golden/*.npz: what the host's FEM / epsilon solve handed to
ibmKernelInterp at every sub-iteration, and the marker forces that came back).

Tolerance: marker forces and fields within relative L2 1e-10 (BASELINE.json north_star); support maps bit-exact.
"""
import numpy as np
import pytest

from tests import cases as K

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", K.EXAMPLES_IBM)
def test_supports_bit_exact(case):
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o))
    ctx.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    cnt, idx, jdx, dirac = ctx.ibm_get_supports()
    assert np.array_equal(cnt, g["s_count"])
    assert np.array_equal(idx, g["s_idx"])
    assert np.array_equal(jdx, g["s_jdx"])
    assert np.array_equal(dirac, g["s_dirac"])      # delta weights too: the kernel rounds every operation like the host
    ctx.close()


def test_supports_at_the_lattice_edge():
    """Markers near / outside the lattice keep only in-range sites (src/IBMNode.cpp:167)."""
    from life_b200 import capi
    g = K.golden("Cylinder")
    o = K.make_oracle(g)
    Dx = o.Dx
    pos = np.array([[0.2 * Dx, 0.3 * Dx], [(o.Nx - 1) * Dx, (o.Ny - 1) * Dx], [-0.4 * Dx, 5.5 * Dx],
                    [10.5 * Dx, 7.5 * Dx], [3.49999 * Dx, 2.50001 * Dx], [-3.0 * Dx, 4 * Dx]])
    n = len(pos)
    ctx = capi.Context(K.life_config(o.params, o))
    ctx.ibm_set_markers(pos, np.zeros((n, 2)), np.ones(n), np.ones(n))
    o.set_markers(pos, np.zeros((n, 2)), np.ones(n), np.ones(n))
    o.find_support()
    for a, b in zip(ctx.ibm_get_supports(), o.supports()):
        assert np.array_equal(a, b)
    ctx.close()


@pytest.mark.parametrize("ordered", [1, 0], ids=["ordered", "atomic"])
@pytest.mark.parametrize("case", K.EXAMPLES_IBM)
def test_fsi_trace_replay(case, ordered):
    """LBM step on the GPU, then for every recorded sub-iteration: set_markers(recorded host state) -> interp -> compare
    the forces with what the reference's ibmKernelInterp returned; spread; finally compare the fields."""
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    cfg = K.life_config(o.params, o, ordered=ordered)
    ctx = capi.Context(cfg)
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
            worst = max(worst, K.rel_l2(force, g["trace_force"][k], floor=1e-6))
            last = g["trace_last"][k]
            k += 1
            if last:
                break
        ctx.ibm_spread()
    assert worst < K.TOL, worst
    st = ctx.download_state()
    for name in ("rho", "u", "f", "force_ibm"):
        err = K.rel_l2(K.sampled(st[name], g), g[name])
        assert err < K.TOL, (case, name, err)
    ctx.close()


def test_ordered_spread_is_bit_exact_and_repeatable():
    """Given identical marker forces, the ordered spread reproduces the oracle's marker-ordered sums bit for bit, twice."""
    from life_b200 import capi
    g = K.golden("Honami")          # 3968 markers, 128 bodies, overlapping supports
    o = K.make_oracle(g)
    n = len(g["m_ds"])
    force = np.stack([np.sin(np.arange(n) * 0.37) * 1e-3, np.cos(np.arange(n) * 0.11) * 2e-3], axis=1)
    o.set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    o.find_support()
    o.set_marker_force(force)
    o.ibm_spread()
    want = o.get("force_ibm")
    ctx = capi.Context(K.life_config(o.params, o, ordered=1))
    K.upload_from_oracle(ctx, o)
    ctx.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    for _ in range(2):
        ctx.ibm_set_forces(force)
        ctx.ibm_spread()
        got = ctx.download_state()["force_ibm"]
        assert np.array_equal(got, want)
    # atomic path: same numbers to rounding
    ctx2 = capi.Context(K.life_config(o.params, o, ordered=0))
    K.upload_from_oracle(ctx2, o)
    ctx2.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    ctx2.ibm_set_forces(force)
    ctx2.ibm_spread()
    assert K.rel_l2(ctx2.download_state()["force_ibm"], want) < 1e-14
    ctx.close()
    ctx2.close()


def test_interp_is_bit_exact_given_the_same_fields():
    """With the oracle's own post-stream populations uploaded, interpolation + forceCalc return the oracle's doubles."""
    from life_b200 import capi
    g = K.golden("Cylinder")
    o = K.make_oracle(g)
    o.step(20)
    o.t = 21
    o.lbm_kernel()
    o.set_markers(g["m_pos"], g["m_vel"] + 0.01, g["m_ds"], g["m_eps"])
    o.find_support()
    # GPU gets the mid-step state: f after lbmKernel, no IBM force yet
    ctx = capi.Context(K.life_config(o.params, o))
    ctx.upload_state(o.get("f"), None, None, o.get("force_xy"), None, o.get("u_in"), o.get("rho_in"))
    ctx.ibm_set_markers(g["m_pos"], g["m_vel"] + 0.01, g["m_ds"], g["m_eps"])
    force = ctx.ibm_interp()
    o.ibm_interp()
    irho, imom = ctx.ibm_get_interp()
    orho, omom = o.interp_values()
    assert np.array_equal(irho, orho)
    assert np.array_equal(imom, omom)
    assert np.array_equal(force, o.marker_force())
    ctx.close()


def test_support_overflow_is_reported():
    """More than 9 sites cannot happen with the 3-point delta on a regular lattice; the error path is exercised by the
    reference's own check (src/IBMNode.cpp:171-172) only in theory.  What must work: zero markers, and re-sizing."""
    from life_b200 import capi
    g = K.golden("Cylinder")
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o))
    K.upload_from_oracle(ctx, o)
    ctx.ibm_set_markers(np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0), np.zeros(0))
    assert ctx.ibm_interp().shape == (0, 2)
    ctx.ibm_spread()
    n = 1000   # grow past the initial capacity
    pos = np.stack([np.linspace(0.3, 1.5, n), np.full(n, 0.2)], axis=1)
    ctx.ibm_set_markers(pos, np.zeros((n, 2)), np.ones(n), np.ones(n))
    ctx.step(1)
    f = ctx.ibm_interp()
    assert f.shape == (n, 2) and np.isfinite(f).all()
    ctx.close()


def _groups(g):
    """marker groups of computeEpsilon: one group with every marker under UNI_EPSILON, else one per body (Objects.cpp:238-251)"""
    n = len(g["m_ds"])
    if int(g["uni_epsilon"]):
        return [np.arange(n)]
    body = g["m_body"]
    return [np.nonzero(body == b)[0] for b in np.unique(body)]


@pytest.mark.parametrize("case", K.EXAMPLES_IBM)
def test_device_epsilon_matches_lapack(case):
    """life_ibm_compute_epsilon (optional device path of computeEpsilon + solveLAPACK) against the epsilon the compiled
    reference obtained from LAPACK: at t = 0 for every body, and along the recorded FSI trace for the flexible ones."""
    from life_b200 import capi
    g = K.golden(case)
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o))
    n = len(g["m_ds"])
    groups = _groups(g)
    ctx.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], np.zeros(n))
    eps = ctx.ibm_compute_epsilon(groups)
    assert K.rel_l2(eps, g["m_eps"]) < K.TOL, K.rel_l2(eps, g["m_eps"])
    # the value is also what the device now spreads with
    _, _, _, _ = ctx.ibm_get_supports()
    # along the trace: the reference recomputes epsilon of flexible bodies (all markers under UNI_EPSILON) every sub-iteration
    flex = g["m_flex"] == 0      # eFlexible = 0 (inc/defs.h:49)
    if int(g["uni_epsilon"]):
        flex = np.ones(n, bool) if flex.any() else flex
    if flex.any():
        sel = [grp for grp in groups if flex[grp].all()]
        worst = 0.0
        for k in range(0, len(g["trace_step"]), max(1, len(g["trace_step"]) // 8)):
            ctx.ibm_set_markers(g["trace_pos"][k], g["trace_vel"][k], g["trace_ds"][k], g["trace_eps"][k])
            eps = ctx.ibm_compute_epsilon(sel)
            worst = max(worst, K.rel_l2(eps, g["trace_eps"][k]))
        assert worst < K.TOL, worst
    ctx.close()


def test_device_epsilon_groups_and_errors():
    from life_b200 import capi
    g = K.golden("Cylinder")
    o = K.make_oracle(g)
    ctx = capi.Context(K.life_config(o.params, o))
    n = len(g["m_ds"])
    start = np.linspace(0.5, 0.9, n)
    ctx.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], start)
    # no groups: nothing changes; a subset: only its members change
    assert np.array_equal(ctx.ibm_compute_epsilon([]), start)
    half = np.arange(0, n, 2)
    eps = ctx.ibm_compute_epsilon([half])
    rest = np.setdiff1d(np.arange(n), half)
    assert np.array_equal(eps[rest], start[rest])
    assert not np.array_equal(eps[half], start[half]) and np.isfinite(eps).all()
    # oracle: same subset as one body
    o.set_markers(g["m_pos"][half], g["m_vel"][half], g["m_ds"][half], start[half])
    o.find_support()
    o.compute_epsilon(0, len(half))
    assert K.rel_l2(eps[half], o.ds_eps()[1]) < K.TOL
    with pytest.raises(capi.LifeError) as e:
        ctx.ibm_compute_epsilon([np.array([0, n + 3])])
    assert e.value.code == capi.E_ARG
    ctx.close()


def test_device_epsilon_matrix_is_bit_exact():
    """life_ibm_assemble_epsilon returns the matrix computeEpsilon builds (src/Objects.cpp:262-301), bit for bit: rebuilt here
    term by term in the reference's order from the oracle's supports and delta function."""
    from life_b200 import capi
    from oracle import oracle as O
    g = K.golden("Cylinder")
    o = K.make_oracle(g)
    n = len(g["m_ds"])
    ctx = capi.Context(K.life_config(o.params, o))
    ctx.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    A = ctx.ibm_assemble_epsilon([np.arange(n)])[0]
    o.set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    o.find_support()
    cnt, idx, jdx, dirac = o.supports()
    dd = O.lib().orc_dirac_delta
    pos, ds, Dx = g["m_pos"], g["m_ds"], o.Dx
    want = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            acc = 0.0
            for s in range(cnt[i]):
                dj = dd(abs(pos[j, 0] / Dx - idx[i, s])) * dd(abs(pos[j, 1] / Dx - jdx[i, s]))
                acc += dirac[i, s] * dj
            want[i, j] = acc * (1.0 * 1.0 * ds[j])
    assert np.array_equal(A, want)
    # and the host's LAPACK on that matrix gives the reference's epsilon
    import scipy.linalg as sl
    eps = sl.solve(A, np.ones(n))
    assert K.rel_l2(eps, g["m_eps"]) < 1e-12
    ctx.close()
