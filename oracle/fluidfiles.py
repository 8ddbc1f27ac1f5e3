"""The reference's two fluid file formats restated in numpy — the ORACLE of the device-fed file paths (SURVEY.md §8f row 2;
include/life_b200.h "device-fed files").  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and tests/mp_parity.py import
this; the product (life_b200/) never does.

  Results/VTK/Fluid.<t>.vti       GridClass::writeVTK      src/Grid.cpp:790-898
      an XML head ending in '_', then three raw blocks — Density, Pressure (Nx*Ny doubles each) and Velocity (3 doubles per node,
      z = 0) — each preceded by its byte count as a UInt64 and each in j-major order (for j: for i:), then the XML tail
  Results/Restart/Fluid.restart   GridClass::writeRestart  src/Grid.cpp:1163-1229, read back by readRestart :1072-1160
      int t, Nx, Ny; double omega, Dx, Dt, Dm; then per node in i-major order int i, j; double rho, ux, uy, force_ibm_x,
      force_ibm_y; double f[9]  -> 44 + 120*Nx*Ny bytes, little endian, written field by field (no padding)

Pinned: tests/test_output_files.py holds vti_bytes() / restart_bytes() to the compiled, unmodified reference's own writers byte
for byte (5 cases x 2 states) and feeds restart_bytes() to the reference's own reader.
"""
import numpy as np


def _g(x):
    """default ostream formatting of a double: 6 significant digits, %g style"""
    return "%g" % x


def vti_frame(Nx, Ny, Dx):
    n8 = Nx * Ny * 8
    head = ('<?xml version="1.0"?>\n'
            '<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n'
            '\t<ImageData WholeExtent="0 %d 0 %d 0 0" Origin="0 0 0" Spacing="%s %s %s">\n'
            '\t\t<Piece Extent="0 %d 0 %d 0 0">\n'
            '\t\t\t<PointData>\n'
            '\t\t\t\t<DataArray type="Float64" Name="Density" format="appended" offset="0"/>\n'
            '\t\t\t\t<DataArray type="Float64" Name="Pressure" format="appended" offset="%d"/>\n'
            '\t\t\t\t<DataArray type="Float64" Name="Velocity" NumberOfComponents="3" format="appended" offset="%d"/>\n'
            '\t\t\t</PointData>\n'
            '\t\t</Piece>\n'
            '\t</ImageData>\n'
            '\t<AppendedData encoding="raw">\n'
            '\t\t_' % (Nx - 1, Ny - 1, _g(Dx), _g(Dx), _g(Dx), Nx - 1, Ny - 1, n8 + 8, 2 * (n8 + 8)))
    tail = "\n\t</AppendedData>\n</VTKFile>\n"
    return head.encode(), tail.encode()


def vti_bytes(rho, u, Dx, Dt, Dm, Drho, rho_p, ref_P=0.0):
    """rho (Nx, Ny), u (Nx, Ny, 2) in lattice units -> the file's bytes.  Operation order of src/Grid.cpp:863, :873, :883-884."""
    rho = np.asarray(rho, np.float64)
    u = np.asarray(u, np.float64)
    Nx, Ny = rho.shape
    c_s = 1.0 / np.sqrt(3.0)
    density = rho * Drho
    pressure = ref_P + (rho - rho_p / Drho) * (c_s * c_s) * Dm / (Dx * (Dt * Dt))
    vel = np.zeros((Nx, Ny, 3))
    vel[:, :, 0] = u[:, :, 0] * (Dx / Dt)
    vel[:, :, 1] = u[:, :, 1] * (Dx / Dt)
    head, tail = vti_frame(Nx, Ny, Dx)
    n8 = np.array([Nx * Ny * 8], "<u8")
    return b"".join([head, n8.tobytes(), np.ascontiguousarray(density.T).tobytes(),
                     n8.tobytes(), np.ascontiguousarray(pressure.T).tobytes(),
                     (3 * n8).tobytes(), np.ascontiguousarray(vel.transpose(1, 0, 2)).tobytes(), tail])


def read_vti(path, Nx, Ny):
    """-> dict(density (Nx,Ny), pressure (Nx,Ny), velocity (Nx,Ny,3)) in the file's physical units"""
    raw = open(path, "rb").read()
    start = raw.index(b"<AppendedData encoding=\"raw\">\n\t\t_") + len(b"<AppendedData encoding=\"raw\">\n\t\t_")
    n = Nx * Ny
    off = start
    out = {}
    for name, comp in (("density", 1), ("pressure", 1), ("velocity", 3)):
        size = int(np.frombuffer(raw, "<u8", 1, off)[0])
        assert size == n * comp * 8, (name, size)
        a = np.frombuffer(raw, "<f8", n * comp, off + 8)
        out[name] = (a.reshape(Ny, Nx).T if comp == 1 else a.reshape(Ny, Nx, 3).transpose(1, 0, 2)).copy()
        off += 8 + size
    assert raw[off:] == b"\n\t</AppendedData>\n</VTKFile>\n"
    return out


# ---- Fluid.restart ---------------------------------------------------------------------------------------------------------
_HEAD = np.dtype([("t", "<i4"), ("Nx", "<i4"), ("Ny", "<i4"), ("omega", "<f8"), ("Dx", "<f8"), ("Dt", "<f8"), ("Dm", "<f8")])
_NODE = np.dtype([("i", "<i4"), ("j", "<i4"), ("rho", "<f8"), ("u", "<f8", (2,)), ("force_ibm", "<f8", (2,)), ("f", "<f8", (9,))])
assert _HEAD.itemsize == 44 and _NODE.itemsize == 120


def read_restart(path):
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


def restart_bytes(t, omega, Dx, Dt, Dm, rho, u, force_ibm, f):
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
