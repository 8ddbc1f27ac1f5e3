"""Timing of the device-fed file paths (life_write_vtk / life_write_restart / life_read_restart) next to the reference's own
writers and reader (compiled reference, oracle/_ref/libref_syn_bgk.so, 4096^2 cavity) on the same box.

    python scripts/io_bench.py [N ...]        default: 4096 8192      -> prints a table; run on a B200 box

For each lattice N x N: the file sizes, synchronous write time, and for the asynchronous mode the time the call blocks the time
loop, how many steps the loop completed while the worker was writing, and the step rate during / outside the write.
Measurement only (reads nothing under oracle/ except through the reference leg, like bench.py's cpu_baseline).
"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from life_b200 import capi  # noqa: E402


def cavity(N, device=0):
    cfg = capi.Config(collision=capi.BGK, Nx=N, Ny=N, omega=1.0, wall_top=capi.VELOCITY, Dx=1.0 / (N - 1), Dt=1.0 / (N - 1) * 0.1,
                      Dm=1.0 / (N - 1) ** 3, Drho=1.0, device=device)
    ctx = capi.Context(cfg)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    cols = max(1, (1 << 27) // (N * 9 * 8))
    u_in = np.zeros((N, 2))
    ctx.upload_begin(u_in, None)
    for c0 in range(0, N, cols):
        nc = min(cols, N - c0)
        ctx.upload_columns(c0, nc, np.broadcast_to(w, (nc, N, 9)).copy())
    ctx.upload_end()
    return ctx


def rate(ctx, t0, n):
    ctx.sync()
    a = time.perf_counter()
    ctx.step_n(t0, n)
    ctx.sync()
    return n / (time.perf_counter() - a)


def ours(N, out):
    ctx = cavity(N)
    ctx.step_n(1, 20)
    t = 21
    base = rate(ctx, t, 50); t += 50
    nodes = N * N
    print("\n== %d x %d (%.1f M nodes), free-running loop %.1f steps/s = %.0f MLUPS" % (N, N, nodes / 1e6, base, base * nodes / 1e6))
    for kind in ("vti", "restart"):
        path = os.path.join(out, "Fluid.%d.vti" % N if kind == "vti" else "Fluid.restart")
        write = (lambda m: ctx.write_vtk(path, 1.0, 0.0, m)) if kind == "vti" else (lambda m: ctx.write_restart(path, t, m))
        write(capi.IO_SYNC)                       # warm: staging buffers allocated
        a = time.perf_counter(); write(capi.IO_SYNC); over_s = time.perf_counter() - a      # overwrites the pages of an existing file
        size = os.path.getsize(path)
        os.remove(path)                           # every Fluid.<t>.vti is a NEW file in a real run: time that
        a = time.perf_counter(); write(capi.IO_SYNC); sync_s = time.perf_counter() - a
        os.remove(path)
        ctx.sync()
        a = time.perf_counter(); write(capi.IO_ASYNC); ctx.sync(); block_s = time.perf_counter() - a     # snapshot included
        done = 0
        a = time.perf_counter()
        while True:
            ctx.step_n(t, 10); t += 10; done += 10
            ctx.sync()
            if not ctx.io_busy():
                break
        during_s = time.perf_counter() - a
        ctx.io_wait()
        job_s, nbytes, was_async = ctx.io_stats()
        print("%-8s %8.1f MB | sync, new file %.3f s (%.2f GB/s), over an existing file %.3f s (%.2f GB/s) | async, new file: call blocks %.1f ms, "
              "worker %.3f s (%.2f GB/s), %d steps done meanwhile at %.1f steps/s (%.0f %% of free-running)%s"
              % (kind, size / 1e6, sync_s, size / sync_s / 1e9, over_s, size / over_s / 1e9, block_s * 1e3, job_s, nbytes / job_s / 1e9, done, done / during_s,
                 100.0 * done / during_s / base, "" if was_async else "  [fell back to sync: snapshot did not fit]"))
    if N <= 8192:
        a = time.perf_counter()
        back = capi.Context(ctx.cfg)
        tb = back.read_restart(os.path.join(out, "Fluid.restart"))
        rd = time.perf_counter() - a
        print("%-8s life_create + life_read_restart %.3f s (%.2f GB/s), t = %d" % ("read", rd, os.path.getsize(os.path.join(out, "Fluid.restart")) / rd / 1e9, tb))
        back.close()
    ctx.close()


def reference(out):
    from oracle import refharness
    if not refharness.available("syn_bgk"):
        print("\n(reference leg skipped: oracle/_ref/libref_syn_bgk.so not built)")
        return
    r = refharness.RefCase("syn_bgk")
    r.step(2)
    a = time.perf_counter(); p = r.write_vtk(); v = time.perf_counter() - a
    vs = os.path.getsize(p)
    a = time.perf_counter(); d = r.write_restart(); w = time.perf_counter() - a
    ws = os.path.getsize(os.path.join(d, "Fluid.restart"))
    a = time.perf_counter(); r.read_restart(); rd = time.perf_counter() - a
    print("\n== reference's own writers / reader, %d x %d, from host arrays (src/Grid.cpp:790, :1163, :1072), files in %s" % (r.Nx, r.Ny, r.workdir))
    print("vti      %8.1f MB | %.3f s (%.3f GB/s)" % (vs / 1e6, v, vs / v / 1e9))
    print("restart  %8.1f MB | write %.3f s (%.3f GB/s) | read %.3f s (%.3f GB/s)   [Fluid.restart + the empty IBM.restart]"
          % (ws / 1e6, w, ws / w / 1e9, rd, ws / rd / 1e9))
    r.close()


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192]
    out = tempfile.mkdtemp(prefix="life_io_")
    print("files under", out, "| free space %.0f GB" % (shutil.disk_usage(out).free / 1e9))
    try:
        for N in sizes:
            ours(N, out)
        reference(out)
    finally:
        shutil.rmtree(out, ignore_errors=True)
