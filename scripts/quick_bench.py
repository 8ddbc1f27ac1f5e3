#!/usr/bin/env python3
"""Developer micro-benchmark: bulk-kernel time per variant / collision operator at one lattice size (1 GPU).

    python scripts/quick_bench.py [N] [steps]

Prints MLUPS and achieved GB/s (144 B per lattice update) from the library's own CUDA-event timing of the bulk kernel,
and the whole-step figure from wall clock around a synchronised loop.  Not the contract benchmark (that is bench.py).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from life_b200 import capi  # noqa: E402

W = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    Dx = 1.0 / (N - 1)
    nu_p = (1.0 / 6.0) / (0.1 * (N - 1))
    Dt = Dx * Dx * (1.0 / 6.0 * 1.0000000000000002) / nu_p
    C = min(N, 1024)
    f = np.empty((C, N, 9))
    f[...] = W
    u_in = np.tile(np.array([[0.1, 0.0]]), (N, 1))
    for coll, cname in ((capi.BGK, "bgk"), (capi.CENTRAL_MOMENTS, "cm")):
        for kern, kname in ((capi.KERNEL_DIRECT, "direct"), (capi.KERNEL_SHUFFLE, "shuffle"), (capi.KERNEL_TMA, "tma")):
            cfg = capi.Config(Nx=N, Ny=N, omega=1.0, collision=coll, wall_top=capi.VELOCITY, Dx=Dx, Dt=Dt, Dm=Dx ** 3,
                              kernel=kern)
            ctx = capi.Context(cfg)
            ctx.upload_begin(u_in, None)
            for il0 in range(0, N, C):
                ctx.upload_columns(il0, min(C, N - il0), f[:min(C, N - il0)])
            ctx.upload_end()
            ctx.step_n(1, 5)
            ctx.sync()
            ctx.set_profiling(True)
            t0 = time.perf_counter()
            ctx.step_n(6, steps)
            ctx.sync()
            wall = (time.perf_counter() - t0) / steps
            ms, n = ctx.bulk_kernel_ms()
            vmax, nan, _, _ = ctx.max_speed()
            mlups_k = N * N / (ms * 1e-3) / 1e6
            print("%-3s %-8s N=%d  bulk %.3f ms  %.0f MLUPS  %.0f GB/s | step(wall) %.3f ms %.0f MLUPS | vmax=%.4f nan=%s"
                  % (cname, kname, N, ms, mlups_k, mlups_k * 144e-3, wall * 1e3, N * N / wall / 1e6, vmax, nan), flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
