"""The oracle (oracle/life_oracle.c) against the golden fixtures the compiled reference wrote (tests/golden/*.npz).

This is the pin of the oracle that travels: the fixtures are committed, /root/reference is not needed.
BGK cases must reproduce the reference bit for bit; central-moments cases to 1e-13 (factored back-transform).
"""
import numpy as np
import pytest

from tests import cases as K


@pytest.mark.parametrize("case", K.EXAMPLES_LBM + K.EXTRA)
def test_lbm_fields_match_reference(case):
    g = K.golden(case)
    o = K.make_oracle(g)
    # initial state (initialiseGrid restatement) at the sampled nodes
    assert np.array_equal(K.sampled(o.get("f"), g), g["init_f"]) or K.rel_l2(K.sampled(o.get("f"), g), g["init_f"]) < 1e-15
    assert np.array_equal(o.get("u_in"), g["u_in"])
    assert np.array_equal(o.get("rho_in"), g["rho_in"])
    assert np.array_equal(o.bcvec(), g["bcvec"])
    o.step(int(g["steps"]))
    exact = not int(g["central_moments"])
    for name in ("f", "rho", "u"):
        mine, ref = K.sampled(o.get(name), g), g[name]
        if exact:
            assert np.array_equal(mine, ref), name
        else:
            assert K.rel_l2(mine, ref) < 1e-13, name
    # checksums over every node
    f = o.get("f")
    tol = 0.0 if exact else 1e-13
    assert abs(o.get("rho").sum() - float(g["sum_rho"])) <= tol * abs(float(g["sum_rho"]))
    assert np.all(np.abs(f.reshape(-1, 9).sum(axis=0) - g["sum_f_per_v"]) <= tol * np.abs(g["sum_f_per_v"]))
    assert abs((o.get("u") ** 2).sum() - float(g["sum_u2"])) <= max(tol, 0.0) * float(g["sum_u2"])


@pytest.mark.parametrize("case", K.EXAMPLES_IBM)
def test_supports_match_reference(case):
    g = K.golden(case)
    o = K.make_oracle(g)
    o.set_markers(g["m_pos"], g["m_vel"], g["m_ds"], g["m_eps"])
    assert o.find_support() == 0
    cnt, idx, jdx, dirac = o.supports()
    assert np.array_equal(cnt, g["s_count"])
    assert np.array_equal(idx, g["s_idx"])
    assert np.array_equal(jdx, g["s_jdx"])
    assert np.array_equal(dirac, g["s_dirac"])


@pytest.mark.parametrize("case", K.EXAMPLES_IBM)
def test_fsi_trace_replay(case):
    """Replay the recorded host side of every sub-iteration (marker pos/vel/ds/epsilon as the reference's FEM and
    epsilon solve produced them) and check what comes back across the seam: marker forces and the final fields."""
    g = K.golden(case)
    o = K.make_oracle(g)
    steps = g["trace_step"]
    k = 0
    for t in range(1, int(g["steps"]) + 1):
        o.t = t
        o.lbm_kernel()
        while True:
            assert steps[k] == t
            o.set_markers(g["trace_pos"][k], g["trace_vel"][k], g["trace_ds"][k], g["trace_eps"][k])
            o.find_support()
            o.ibm_interp()
            assert np.array_equal(o.marker_force(), g["trace_force"][k]), (t, k)
            last = g["trace_last"][k]
            k += 1
            if last:
                break
        o.ibm_spread()
    assert k == len(steps)
    for name in ("f", "rho", "u", "force_ibm"):
        assert np.array_equal(K.sampled(o.get(name), g), g[name]), name


def test_rigid_epsilon_and_ds():
    """computeDs / computeEpsilon restatement (plain LU instead of LAPACK) against the reference's values."""
    g = K.golden("Cylinder")
    o = K.make_oracle(g)
    n = len(g["m_ds"])
    o.set_markers(g["m_pos"], g["m_vel"], np.zeros(n), np.zeros(n))
    o.find_support()
    o.compute_ds(0, n)
    o.compute_epsilon(0, n)
    ds, eps = o.ds_eps()
    assert np.array_equal(ds, g["m_ds"])
    assert K.rel_l2(eps, g["m_eps"]) < 1e-12
