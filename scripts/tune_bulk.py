#!/usr/bin/env python3
"""Sweep over cfg.tune (CTA size / cache hints / occupancy of the bulk sweep) and the kernel variants at one lattice size, next to
the memory-system ceilings of the same box (life_membw).  python scripts/tune_bulk.py [N] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from life_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C = 1024
f = np.empty((C, N, 9))
f[...] = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
u_in = np.tile(np.array([[0.1, 0.0]]), (N, 1))
Dx = 1.0 / (N - 1)
names = {0: "read only", 1: "write only", 2: "copy LDG/STG 16 B", 3: "copy LDG/STG 32 B", 4: "copy TMA bulk g2s + s2g", 5: "copy 9 planes -> 9 planes, 16 B, sweep launch shape", 6: "copy 9 planes -> 9 planes, 32 B"}
for mode in range(7):
    print("membw mode %d (%s): %.0f GB/s" % (mode, names[mode], capi.membw(mode, 16 << 30, 5)), flush=True)
for coll, cname in ((capi.BGK, "bgk"), (capi.CENTRAL_MOMENTS, "cm")):
    for tune, kernel in ((0, 0), (1, 0), (4, 0), (0, 4), (1, 4), (0, 3), (0, 1), (0, 0)):
        cfg = capi.Config(Nx=N, Ny=N, omega=1.0, collision=coll, wall_top=capi.VELOCITY, Dx=Dx, Dt=Dx * Dx * 0.5, Dm=Dx ** 3, tune=tune, kernel=kernel)
        ctx = capi.Context(cfg)
        ctx.upload_begin(u_in, None)
        for il0 in range(0, N, C):
            ctx.upload_columns(il0, min(C, N - il0), f[:min(C, N - il0)])
        ctx.upload_end()
        ctx.step_n(1, 5)
        ctx.sync()
        best = 1e9
        for rep in range(3):
            ctx.set_profiling(True)
            ctx.step_n(6 + rep * steps, steps)
            ms, n = ctx.bulk_kernel_ms()
            best = min(best, ms)
        print("%s kernel=%d tune=%2d  bulk %.4f ms  %.0f MLUPS  %.0f GB/s" % (cname, kernel, tune, best, N * N / best / 1e3, N * N * 144 / best / 1e6), flush=True)
        ctx.close()

# ---- force modes of the sweep: uniform force_xy, IBM force planes (with / without the span mask), both -------------------
def timed(ctx, t0):
    best = 1e9
    for rep in range(3):
        ctx.set_profiling(True)
        ctx.step_n(t0 + rep * steps, steps)
        ms, n = ctx.bulk_kernel_ms()
        best = min(best, ms)
    return best

n_mk = 8192
th = 2.0 * np.pi * np.arange(n_mk) / n_mk
for coll, cname in ((capi.BGK, "bgk"), (capi.CENTRAL_MOMENTS, "cm")):
    for label, uni, ibm, tune in (("uniform force_xy", 1, 0, 0), ("force_ibm (8192 markers), span mask", 0, 1, 0),
                                  ("force_ibm, planes read everywhere (round 1)", 0, 1, 20), ("uniform + force_ibm, span mask", 1, 1, 0)):
        cfg = capi.Config(Nx=N, Ny=N, omega=1.0, collision=coll, wall_top=capi.VELOCITY, Dx=Dx, Dt=Dx * Dx * 0.5, Dm=Dx ** 3, tune=tune, ordered=1)
        ctx = capi.Context(cfg)
        ctx.upload_begin(u_in, None)
        fxy = np.zeros((C, N, 2)); fxy[..., 0] = 1e-7
        for il0 in range(0, N, C):
            ctx.upload_columns(il0, C, f[:C], None, None, fxy if uni else None, None)
        ctx.upload_end()
        ctx.step_n(1, 3)
        if ibm:
            # a ring of markers in the middle of the lattice (physical units: the lattice spans [0, 1])
            pos = np.stack([0.5 + 0.2 * np.cos(th), 0.5 + 0.2 * np.sin(th)], axis=1)
            ctx.ibm_set_markers(pos, np.zeros((n_mk, 2)), np.full(n_mk, 1.0), np.full(n_mk, 1.0))
            ctx.ibm_interp()
            ctx.ibm_spread()
        best = timed(ctx, 4)
        print("%s %-48s bulk %.4f ms  %.0f MLUPS  %.0f GB/s of the 144 B/node" % (cname, label, best, N * N / best / 1e3, N * N * 144 / best / 1e6), flush=True)
        ctx.close()
