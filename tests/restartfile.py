"""Readers for the reference's binary restart files — the parity dump format of its own regression protocol
(testing/run-tests.sh:100 diffs Results/ against RefData/Results/).

Fluid.restart (src/Grid.cpp:1183-1221): int t, Nx, Ny; double omega, Dx, Dt, Dm; then per node in i-major order
    int i, j; double rho, ux, uy, force_ibm_x, force_ibm_y; double f[9]            -> 44 + 120*Nx*Ny bytes
IBM.restart (src/Objects.cpp:1226-1257): size_t nBodies; per body { int ID; size_t nNodes;
    per node { double posx, posy, velx, vely, forcex, forcey } }
Little endian, native int / size_t / double, written field by field (no padding).
"""
import numpy as np

from oracle.fluidfiles import read_restart as read_fluid  # noqa: F401  (format restatement: oracle/fluidfiles.py)
from oracle.fluidfiles import restart_bytes as fluid_bytes  # noqa: F401


def read_ibm(path):
    raw = open(path, "rb").read()
    off = 0
    nb = int(np.frombuffer(raw, "<u8", 1, off)[0]); off += 8
    bodies = []
    for _ in range(nb):
        bid = int(np.frombuffer(raw, "<i4", 1, off)[0]); off += 4
        nn = int(np.frombuffer(raw, "<u8", 1, off)[0]); off += 8
        a = np.frombuffer(raw, "<f8", 6 * nn, off).reshape(nn, 6); off += 48 * nn
        bodies.append(dict(id=bid, pos=a[:, 0:2].copy(), vel=a[:, 2:4].copy(), force=a[:, 4:6].copy()))
    assert off == len(raw)
    return bodies


def read_table(path):
    """TotalForces.out / TipPositions.out: whitespace-separated numeric rows (header lines skipped)."""
    rows = []
    for ln in open(path):
        p = ln.split()
        try:
            rows.append([float(x) for x in p])
        except ValueError:
            continue
    w = max(len(r) for r in rows)
    return np.array([r for r in rows if len(r) == w])
