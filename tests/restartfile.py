"""Readers for the reference's binary restart files — the parity dump format of its own regression protocol
(testing/run-tests.sh:100 diffs Results/ against RefData/Results/).

Fluid.restart (src/Grid.cpp:1183-1221): int t, Nx, Ny; double omega, Dx, Dt, Dm; then per node in i-major order
    int i, j; double rho, ux, uy, force_ibm_x, force_ibm_y; double f[9]            -> 44 + 120*Nx*Ny bytes
IBM.restart (src/Objects.cpp:1226-1257): size_t nBodies; per body { int ID; size_t nNodes;
    per node { double posx, posy, velx, vely, forcex, forcey } }
Little endian, native int / size_t / double, written field by field (no padding).
"""
import numpy as np

_HEAD = np.dtype([("t", "<i4"), ("Nx", "<i4"), ("Ny", "<i4"), ("omega", "<f8"), ("Dx", "<f8"), ("Dt", "<f8"), ("Dm", "<f8")])
_NODE = np.dtype([("i", "<i4"), ("j", "<i4"), ("rho", "<f8"), ("u", "<f8", (2,)), ("force_ibm", "<f8", (2,)), ("f", "<f8", (9,))])
assert _HEAD.itemsize == 44 and _NODE.itemsize == 120


def read_fluid(path):
    raw = np.fromfile(path, dtype=np.uint8)
    head = raw[:44].view(_HEAD)[0]
    Nx, Ny = int(head["Nx"]), int(head["Ny"])
    assert raw.size == 44 + 120 * Nx * Ny, (raw.size, Nx, Ny)
    nodes = raw[44:].view(_NODE)
    ii, jj = np.divmod(np.arange(Nx * Ny), Ny)
    assert np.array_equal(nodes["i"], ii) and np.array_equal(nodes["j"], jj)
    out = {k: head[k].item() for k in _HEAD.names}
    out["rho"] = nodes["rho"].reshape(Nx, Ny).copy()
    out["u"] = nodes["u"].reshape(Nx, Ny, 2).copy()
    out["force_ibm"] = nodes["force_ibm"].reshape(Nx, Ny, 2).copy()
    out["f"] = nodes["f"].reshape(Nx, Ny, 9).copy()
    return out


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


def fluid_bytes(t, omega, Dx, Dt, Dm, rho, u, force_ibm, f):
    """The bytes GridClass::writeRestart (src/Grid.cpp:1163-1221) produces for this state: arrays shaped (Nx, Ny[, k])."""
    Nx, Ny = rho.shape
    head = np.zeros(1, _HEAD)
    head["t"], head["Nx"], head["Ny"] = t, Nx, Ny
    head["omega"], head["Dx"], head["Dt"], head["Dm"] = omega, Dx, Dt, Dm
    nodes = np.zeros(Nx * Ny, _NODE)
    nodes["i"], nodes["j"] = np.divmod(np.arange(Nx * Ny), Ny)
    nodes["rho"] = np.asarray(rho).reshape(-1)
    nodes["u"] = np.asarray(u).reshape(-1, 2)
    nodes["force_ibm"] = np.asarray(force_ibm).reshape(-1, 2)
    nodes["f"] = np.asarray(f).reshape(-1, 9)
    return head.tobytes() + nodes.tobytes()
