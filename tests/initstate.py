"""Deterministic non-uniform initial states for the parity tests (no RNG: smooth waves, so every run and every box
builds the same doubles).  Used by tests/golden/make_golden.py to overwrite the reference's uniform initial state and by
the tests to give the oracle and the CUDA library the same start."""
import numpy as np

CX = np.array([0, 1, -1, 0, 0, 1, -1, 1, -1], dtype=np.float64)
CY = np.array([0, 0, 0, 1, -1, 1, -1, -1, 1], dtype=np.float64)
W = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4, dtype=np.float64)


def equilibrium(rho, ux, uy, central_moments):
    """f_eq of src/Grid.cpp:249-264 for arrays rho, ux, uy -> [..., 9]."""
    rho, ux, uy = rho[..., None], ux[..., None], uy[..., None]
    if central_moments:
        return 0.25 * rho * W * (9.0 * CX * CX * ux * ux + 6.0 * CX * ux - 3.0 * ux * ux + 2.0) * \
            (9.0 * CY * CY * uy * uy + 6.0 * CY * uy - 3.0 * uy * uy + 2.0)
    return rho * W * (1.0 + 3.0 * (CX * ux + CY * uy) + 4.5 * (ux * ux * (CX * CX - 1.0 / 3.0) + uy * uy * (CY * CY - 1.0 / 3.0))
                      + 9.0 * CX * CY * ux * uy)


def wavy_state(Nx, Ny, central_moments, amp=0.04, non_equilibrium=0.02):
    """rho, u smooth and periodic in both directions; f = f_eq * (1 + small direction-dependent modulation)."""
    i = np.arange(Nx, dtype=np.float64)[:, None]
    j = np.arange(Ny, dtype=np.float64)[None, :]
    a, b = 2.0 * np.pi * i / Nx, 2.0 * np.pi * j / Ny
    rho = 1.0 + 0.03 * np.cos(a + 2.0 * b) + 0.01 * np.sin(3.0 * a)
    ux = amp * np.sin(b) * np.cos(2.0 * a) + 0.5 * amp
    uy = -amp * np.cos(b + 0.3) * np.sin(a) - 0.25 * amp
    f = equilibrium(rho, ux, uy, central_moments)
    mod = 1.0 + non_equilibrium * np.sin(a[..., None] * (1.0 + np.arange(9) % 3) + b[..., None] * (1.0 + np.arange(9) % 2) + np.arange(9))
    f = f * mod
    # macroscopics consistent with f (what the reference would hold after a completed, unforced step)
    rho = f.sum(axis=-1)
    u = np.stack([(f * CX).sum(axis=-1) / rho, (f * CY).sum(axis=-1) / rho], axis=-1)
    return np.ascontiguousarray(f), np.ascontiguousarray(rho), np.ascontiguousarray(u)
