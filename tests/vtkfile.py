"""The reference's fluid VTK file (Results/VTK/Fluid.<t>.vti, src/Grid.cpp:790-898) restated in numpy: test infrastructure that
turns a (rho, u) state into the exact bytes GridClass::writeVTK would write, and parses such a file back.

Layout: an XML head ending in '_', then three raw blocks — Density, Pressure (Nx*Ny doubles each) and Velocity (3 doubles per
node, z = 0) — each preceded by its byte count as a UInt64 and each in j-major order (for j: for i:), then the XML tail.
tests/test_output_files.py pins this restatement against files written by the compiled reference itself.
"""
import numpy as np


def _g(x):
    """default ostream formatting of a double: 6 significant digits, %g style"""
    return "%g" % x


def frame(Nx, Ny, Dx):
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


def fluid_bytes(rho, u, Dx, Dt, Dm, Drho, rho_p, ref_P=0.0):
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
    head, tail = frame(Nx, Ny, Dx)
    n8 = np.array([Nx * Ny * 8], "<u8")
    return b"".join([head, n8.tobytes(), np.ascontiguousarray(density.T).tobytes(),
                     n8.tobytes(), np.ascontiguousarray(pressure.T).tobytes(),
                     (3 * n8).tobytes(), np.ascontiguousarray(vel.transpose(1, 0, 2)).tobytes(), tail])


def read_fluid(path, Nx, Ny):
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
