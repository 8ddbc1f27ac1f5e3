"""A golden fixture (tests/golden/<case>.npz, written by the compiled reference) turned into everything a run through the
C ABI needs — configuration and initial state — WITHOUT the oracle: the fixture carries the case description and the
scalings the reference's constructor computed (Dx, Dt, Dm, Drho), the inlet arrays, the uniform initial force_xy, and for the
"wavy" cases the parameters of the deterministic initial state (tests/initstate.py), so nothing under oracle/ is imported.
bench.py uses this for its untimed multi-slab parity spot check; tests may use it too."""
import os

import numpy as np

from tests.initstate import wavy_state

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(case):
    return np.load(os.path.join(GOLDEN_DIR, case + ".npz"))


def config_kwargs(g):
    """keyword arguments of life_b200.capi.Config for the fixture's compile-time case"""
    w = g["walls"]
    return dict(collision=int(g["central_moments"]), Nx=int(g["Nx"]), Ny=int(g["Ny"]), omega=float(g["omega"]),
                wall_left=int(w[0]), wall_right=int(w[1]), wall_bottom=int(w[2]), wall_top=int(w[3]),
                inlet_ramp=float(g["inlet_ramp"]), Dx=float(g["Dx"]), Dt=float(g["Dt"]), Dm=float(g["Dm"]), Drho=float(g["Drho"]),
                womersley=float(g["womersley"]), height_p=float(g["height_p"]), nu_p=float(g["nu_p"]),
                gravity_x=float(g["gravityX"]), gravity_y=float(g["gravityY"]), dpdx=float(g["dpdx"]), dpdy=float(g["dpdy"]),
                ordered=int(g["ordered"]))


def initial_state(g):
    """(f, rho, u, force_xy, u_in, rho_in) the golden run started from.  Only for fixtures whose initial state is the
    deterministic wavy one (the others start from initialiseGrid's state, which only the oracle / reference rebuild)."""
    if not int(g["wavy"]):
        raise ValueError("this fixture starts from initialiseGrid's state; it cannot be rebuilt without the oracle")
    Nx, Ny = int(g["Nx"]), int(g["Ny"])
    f, rho, u = wavy_state(Nx, Ny, bool(int(g["central_moments"])), amp=float(g["wavy_amp"]), non_equilibrium=float(g["wavy_neq"]))
    fxy = np.empty((Nx, Ny, 2))
    fxy[...] = np.asarray(g["init_force_xy0"], np.float64).reshape(2)
    return f, rho, u, fxy, np.ascontiguousarray(g["u_in"], np.float64), np.ascontiguousarray(g["rho_in"], np.float64)


def sample_index(g, i_begin, i_end):
    """(mask over the fixture's samples that fall into columns [i_begin, i_end), local i, j of those)"""
    s = g["sample"]
    Ny = int(g["Ny"])
    i, j = s // Ny, s % Ny
    m = (i >= i_begin) & (i < i_end)
    return m, i[m] - i_begin, j[m]
