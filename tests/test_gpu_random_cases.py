"""Seeded random case definitions against the oracle: every wall-type combination the reference's params.h can express
(periodic / wall / velocity / free-slip / pressure on any side, also asymmetric; convective on the right), both collision
operators, random omega, inlet ramp, body force, non-uniform start — on small odd-shaped lattices, where corners, two-node-deep
boundary stencils and wrap-around meet.  Plus two mid-size long runs.  Tolerance: relative L2 1e-10 (north star)."""
import numpy as np
import pytest

from tests import cases as K

pytestmark = pytest.mark.gpu

WALLS = [0, 1, 2, 3, 4]          # eFluid (periodic), eWall, eVelocity, eFreeSlip, ePressure (inc/defs.h:52)


def _random_params(seed):
    from oracle import oracle as O
    r = np.random.RandomState(1000 + seed)
    Nx, Ny = int(r.randint(9, 70)), int(r.randint(9, 70))
    left, bottom, top = (int(r.choice(WALLS)) for _ in range(3))
    right = int(r.choice(WALLS + [5]))                      # eConvective only on the right (src/Grid.cpp:919-922)
    if right == 5 and Nx < 12:
        Nx = 12
    p = O.Params(Nx=Nx, Ny=Ny, central_moments=int(r.randint(2)), omega=float(r.uniform(0.6, 1.9)),
                 wall_left=left, wall_right=right, wall_bottom=bottom, wall_top=top,
                 inlet_ramp=float(r.choice([-1.0, 0.002, 5.0])), profile=int(r.choice([-1, 0, 1, 2])),
                 height_p=1.0, rho_p=float(r.choice([1.0, 1000.0])), nu_p=float(r.uniform(0.002, 0.02)),
                 ux0_p=float(r.uniform(-0.3, 0.3)), uy0_p=float(r.uniform(-0.2, 0.2)),
                 gravityX=float(r.choice([0.0, 0.3])), gravityY=float(r.choice([0.0, -0.2])),
                 dpdx=float(r.choice([0.0, 0.5])), dpdy=0.0,
                 uxInlet_p=float(r.uniform(0.2, 1.0)), uyInlet_p=float(r.choice([0.0, 0.1])))
    return p, r


@pytest.mark.parametrize("kernel", [2, 3], ids=["shuffle", "tma"])
@pytest.mark.parametrize("seed", range(32))
def test_random_case_matches_oracle(seed, kernel):
    from life_b200 import capi
    from oracle import oracle as O
    from tests.initstate import wavy_state
    p, r = _random_params(seed)
    o = O.Oracle(p)
    # the reference stops with ERROR for a corner whose two axis neighbours are fluid (src/Grid.cpp:527-528): so must life_create
    bad_corner = any(o.normal(*divmod(int(b), o.Ny))[2] < 0 for b in o.bcvec())
    cfg = K.life_config(p, o, kernel=kernel)
    if bad_corner:
        with pytest.raises(capi.LifeError) as e:
            capi.Context(cfg)
        assert e.value.code == capi.E_ARG and "Corner" in str(e.value)
        return
    # non-uniform, slightly off-equilibrium start; modest speeds keep every boundary type well-posed for a few dozen steps
    f0, rho0, u0 = wavy_state(o.Nx, o.Ny, bool(p.central_moments), amp=0.03, non_equilibrium=0.01)
    o.set("f", f0)
    o.set("rho", rho0)
    o.set("u", u0)
    ctx = capi.Context(cfg)
    K.upload_from_oracle(ctx, o)
    steps = 25
    for t in range(1, steps + 1):
        ctx.step(t)
    o.step(steps)
    st = ctx.download_state()
    ctx.close()
    ref = {n: o.get(n) for n in ("rho", "u", "f")}
    if not all(np.isfinite(ref[n]).all() for n in ref) or np.abs(ref["u"]).max() > 1.0:
        # e.g. free-slip inlet + convective outlet: the run diverges exponentially in the reference itself, so the 1e-16
        # differences of the first steps (checked below at 5 steps) are amplified without bound
        o2 = O.Oracle(p)
        o2.set("f", f0); o2.set("rho", rho0); o2.set("u", u0)
        ctx = capi.Context(cfg)
        K.upload_from_oracle(ctx, o2)
        for t in range(1, 4):
            ctx.step(t)
        o2.step(3)
        st = ctx.download_state()
        ctx.close()
        for name in ("rho", "u", "f"):
            assert K.rel_l2(st[name], o2.get(name), floor=1e-5 if name == "u" else 0.0) < K.TOL, (seed, "3 steps", name)
        pytest.skip("this random combination diverges in the reference arithmetic itself (matched for the first 3 steps)")
    for name in ("rho", "u", "f"):
        err = K.rel_l2(st[name], ref[name], floor=1e-5 if name == "u" else 0.0)
        assert err < K.TOL, (seed, name, err, p.as_dict())


@pytest.mark.parametrize("walls,cm", [((1, 1, 1, 2), 1), ((2, 4, 1, 1), 0), ((0, 0, 1, 3), 0)],
                         ids=["cavity-cm", "channel-bgk", "periodic-freeslip-bgk"])
def test_mid_size_long_run(walls, cm):
    """1024 x 768 nodes, 400 steps: the same comparison at a size where the lattice no longer fits L2 and every tile / warp
    boundary of the sweep is crossed many times."""
    from life_b200 import capi
    from oracle import oracle as O
    p = O.Params(Nx=1024, Ny=768, central_moments=cm, omega=1.6, wall_left=walls[0], wall_right=walls[1], wall_bottom=walls[2],
                 wall_top=walls[3], nu_p=0.002, uxInlet_p=0.6, profile=0 if walls[0] == 2 else -1,
                 dpdx=0.4 if walls[0] == 0 else 0.0)
    o = O.Oracle(p)
    ctx = capi.Context(K.life_config(p, o))
    K.upload_from_oracle(ctx, o)
    ctx.step_n(1, 400)
    o.step(400)
    st = ctx.download_state()
    ctx.close()
    for name in ("rho", "u", "f"):
        err = K.rel_l2(st[name], o.get(name), floor=1e-5 if name == "u" else 0.0)
        assert err < K.TOL, (name, err)
    assert np.abs(o.get("u")).max() > 1e-4
