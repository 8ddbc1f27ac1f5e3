#!/usr/bin/env python3
"""32768^2 on ONE B200 with the in-place layout (cfg.inplace: 77 GB of populations instead of 155 GB): sweep rate, and an ASYNCHRONOUS
snapshot that now fits beside the lattice (with two buffers, 155 GB, not even the 26 GB .vti snapshot does, and the write
falls back to synchronous).  python scripts/inplace_big.py [N] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from life_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C = 512
f = np.empty((C, N, 9))
f[...] = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
u_in = np.tile(np.array([[0.1, 0.0]]), (N, 1))
Dx = 1.0 / (N - 1)
free0 = torch.cuda.mem_get_info()[0]
ctx = capi.Context(capi.Config(Nx=N, Ny=N, omega=1.0, wall_top=capi.VELOCITY, Dx=Dx, Dt=Dx * Dx * 0.5, Dm=Dx ** 3, inplace=1))
ctx.upload_begin(u_in, None)
for il0 in range(0, N, C):
    ctx.upload_columns(il0, C, f)
ctx.upload_end()
print("lattice %d^2, in place: %.1f GB of device memory in use" % (N, (free0 - torch.cuda.mem_get_info()[0]) / 1e9), flush=True)
ctx.step_n(1, 3)
ctx.sync()
ctx.set_profiling(True)
t0 = time.perf_counter()
ctx.step_n(4, steps)
ctx.sync()
dt = time.perf_counter() - t0
ms, n = ctx.bulk_kernel_ms()
print("in-place sweep: %.3f ms per step (bulk kernel %.3f ms) = %.0f MLUPS, %.0f GB/s of the 144 B/node" % (dt / steps * 1e3, ms, N * N * steps / dt / 1e6, N * N * 144 / ms / 1e6), flush=True)
free = torch.cuda.mem_get_info()[0]
print("free device memory beside the lattice: %.1f GB; an asynchronous .vti snapshot needs %.1f GB (24 B/node), a restart snapshot %.1f GB (96 B/node): %s"
      % (free / 1e9, 24 * N * N / 1e9, 96 * N * N / 1e9, "both fit" if free > 96 * N * N else ("the .vti snapshot fits" if free > 24 * N * N else "neither fits")), flush=True)
vmax, nan, _, _ = ctx.max_speed()
assert not nan and 0 < vmax < 0.2
ctx.close()
