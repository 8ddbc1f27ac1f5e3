"""Shared helpers of the test suite: golden fixtures -> oracle parameters / C-ABI configuration, comparison metrics."""
import os

import numpy as np

from oracle import oracle as O
from tests.initstate import wavy_state

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EXAMPLES_LBM = ["LidDrivenCavity", "ChannelFlow"]                      # no bodies
EXAMPLES_IBM = ["Cylinder", "TurekHron", "InvertedFlag", "Honami", "PELskin"]
EXTRA = ["t_periodic_bgk", "t_periodic_cm", "t_convective", "t_freeslip_cm", "t_womersley", "t_velocity_box",
         "t_pressure_left", "t_yperiodic"]
ALL_CASES = EXAMPLES_LBM + EXAMPLES_IBM + EXTRA

# tolerance of BASELINE.json's north_star: macroscopic fields and marker forces within relative L2 1e-10
TOL = 1e-10


def golden(case):
    return np.load(os.path.join(GOLDEN_DIR, case + ".npz"))


def oracle_params(g):
    """orc_params of a golden fixture (= the compile-time case of the reference build that wrote it)."""
    w = g["walls"]
    return O.Params(Nx=int(g["Nx"]), Ny=int(g["Ny"]), central_moments=int(g["central_moments"]),
                    ordered=int(g["ordered"]), uni_epsilon=int(g["uni_epsilon"]), profile=int(g["profile"]),
                    wall_left=int(w[0]), wall_right=int(w[1]), wall_bottom=int(w[2]), wall_top=int(w[3]),
                    inlet_ramp=float(g["inlet_ramp"]), womersley=float(g["womersley"]), omega=float(g["omega"]),
                    height_p=float(g["height_p"]), rho_p=float(g["rho_p"]), nu_p=float(g["nu_p"]),
                    ux0_p=float(g["ux0_p"]), uy0_p=float(g["uy0_p"]), gravityX=float(g["gravityX"]),
                    gravityY=float(g["gravityY"]), dpdx=float(g["dpdx"]), dpdy=float(g["dpdy"]),
                    uxInlet_p=float(g["uxInlet_p"]), uyInlet_p=float(g["uyInlet_p"]))


def make_oracle(g):
    """Oracle lattice in the state the golden run started from."""
    p = oracle_params(g)
    o = O.Oracle(p)
    if int(g["wavy"]):
        f0, rho0, u0 = wavy_state(o.Nx, o.Ny, bool(p.central_moments), amp=float(g["wavy_amp"]),
                                  non_equilibrium=float(g["wavy_neq"]))
        o.set("f", f0)
        o.set("rho", rho0)
        o.set("u", u0)
    return o


def life_config(p, o, **kw):
    """life_config (C ABI) for oracle parameters `p`; scalings come from the oracle's constructor restatement."""
    from life_b200 import capi
    cfg = capi.Config(collision=capi.CENTRAL_MOMENTS if p.central_moments else capi.BGK, Nx=p.Nx, Ny=p.Ny,
                      omega=p.omega, wall_left=p.wall_left, wall_right=p.wall_right, wall_bottom=p.wall_bottom,
                      wall_top=p.wall_top, inlet_ramp=p.inlet_ramp, Dx=o.Dx, Dt=o.Dt, Dm=o.Dm, Drho=o.Drho,
                      womersley=p.womersley, height_p=p.height_p, nu_p=p.nu_p, gravity_x=p.gravityX,
                      gravity_y=p.gravityY, dpdx=p.dpdx, dpdy=p.dpdy, ordered=p.ordered)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def upload_from_oracle(ctx, o):
    """Hand the oracle's current state to the CUDA library exactly as a LIFE host would after initialiseGrid."""
    ctx.upload_state(o.get("f"), o.get("rho"), o.get("u"), o.get("force_xy"), o.get("force_ibm"), o.get("u_in"),
                     o.get("rho_in"))


def rel_l2(a, b, floor=0.0):
    """||a - b||_2 / ||b||_2.  `floor` (per entry) guards the denominator for quantities that are legitimately ~0:
    populations are O(1), so anything derived from them carries absolute rounding of ~1e-16 per entry and a reference
    norm below floor*sqrt(n) is compared against that floor instead of against itself."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    n = max(np.linalg.norm(b), floor * np.sqrt(max(b.size, 1)))
    d = np.linalg.norm(a - b)
    return d / n if n > 0 else d


def sampled(arr, g):
    s = g["sample"]
    Ny = int(g["Ny"])
    return arr[s // Ny, s % Ny]
