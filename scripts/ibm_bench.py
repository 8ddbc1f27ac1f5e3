#!/usr/bin/env python3
"""Marker-kernel throughput (markers/s for gather and scatter; latency-bound, no roofline claim).

    python scripts/ibm_bench.py [N] [reps]

Synthetic markers: circles of ~1-lattice-unit spacing scattered over an N x N BGK lattice (default 4096), marker counts from
the examples' scale (62, 3968) to 1M.  Times, per call and synchronised like the host program uses them:
  update+gather : life_ibm_set_markers (one H2D copy + support search) + life_ibm_interp (gather, force, D2H of the forces)
  scatter       : life_ibm_spread, ordered (cell-list gather, bit-repeatable) and atomic
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from life_b200 import capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
Dx = 1.0 / (N - 1)
f = np.empty((N, N, 9))
f[...] = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
u_in = np.tile(np.array([[0.1, 0.0]]), (N, 1))


def markers(n):
    """rings of 64 markers with unit spacing (radius 64/(2 pi) lattice units), centres on a jittered grid"""
    rings = max(1, n // 64)
    per = n // rings
    side = int(np.ceil(np.sqrt(rings)))
    pitch = (N - 80) / side
    k = np.arange(rings)
    cx = 40 + (k % side + 0.5) * pitch + 0.37 * np.sin(k)
    cy = 40 + (k // side + 0.5) * pitch + 0.41 * np.cos(k)
    th = 2 * np.pi * np.arange(per) / per
    r = min(per / (2 * np.pi), pitch / 2 - 2)
    x = (cx[:, None] + r * np.cos(th)[None, :]).ravel()
    y = (cy[:, None] + r * np.sin(th)[None, :]).ravel()
    return np.stack([x, y], axis=1) * Dx


print("lattice %d x %d, %d repetitions per figure" % (N, N, reps))
for ordered in (1, 0):
    cfg = capi.Config(Nx=N, Ny=N, omega=1.0, wall_top=capi.VELOCITY, Dx=Dx, Dt=Dx * Dx * 0.5, Dm=Dx ** 3, ordered=ordered)
    ctx = capi.Context(cfg)
    ctx.upload_state(f, None, None, None, None, u_in, None)
    ctx.step_n(1, 3)
    for n in (62, 3968, 65536, 1048576):
        pos = markers(n)
        n = len(pos)
        vel = np.zeros((n, 2))
        ds = np.ones(n)
        eps = np.full(n, 2.0)
        ctx.ibm_set_markers(pos, vel, ds, eps)
        ctx.ibm_interp()
        ctx.ibm_spread()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.ibm_set_markers(pos, vel, ds, eps)
            ctx.ibm_interp()
        tg = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.ibm_spread()
        ctx.sync()
        ts = (time.perf_counter() - t0) / reps
        print("%-7s n=%8d  update+gather %9.1f us (%8.1f M markers/s)   scatter %9.1f us (%8.1f M markers/s)"
              % ("ordered" if ordered else "atomic", n, tg * 1e6, n / tg / 1e6, ts * 1e6, n / ts / 1e6), flush=True)
    ctx.close()
