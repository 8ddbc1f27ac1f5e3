#!/usr/bin/env python3
"""Developer sweep over cfg.tune (CTA size / cache hints of the bulk sweep) at one lattice size.  python scripts/tune_bulk.py [N] [steps]"""
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
for coll, cname in ((capi.BGK, "bgk"), (capi.CENTRAL_MOMENTS, "cm")):
    for tune in (0, 1, 3, 11, 12, 13):
        cfg = capi.Config(Nx=N, Ny=N, omega=1.0, collision=coll, wall_top=capi.VELOCITY, Dx=Dx, Dt=Dx * Dx * 0.5, Dm=Dx ** 3, tune=tune)
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
        print("%s tune=%2d  bulk %.4f ms  %.0f MLUPS  %.0f GB/s" % (cname, tune, best, N * N / best / 1e3, N * N * 144 / best / 1e6), flush=True)
        ctx.close()
