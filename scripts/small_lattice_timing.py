#!/usr/bin/env python3
"""Small-lattice regime (BASELINE configs[0..1]): device time per time step of the body-free examples, per-step launches (cfg.tune = 30)
against the persistent cluster kernel behind life_step_n (csrc/lbm_small.cu).  python scripts/small_lattice_timing.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from life_b200 import capi  # noqa: E402
from tests import fixture_state as FS  # noqa: E402
from tests.initstate import equilibrium  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
for case in ("LidDrivenCavity", "ChannelFlow", "t_periodic_cm", "t_womersley"):
    g = FS.load(case)
    kw = FS.config_kwargs(g)
    Nx, Ny = kw["Nx"], kw["Ny"]
    if int(g["wavy"]):
        f, rho, u, fxy, u_in, rho_in = FS.initial_state(g)
    else:   # rest state + the fixture's inlet arrays (timing only)
        rho, u = np.ones((Nx, Ny)), np.zeros((Nx, Ny, 2))
        f = equilibrium(rho, u[..., 0], u[..., 1], bool(kw["collision"]))
        fxy, u_in, rho_in = None, np.ascontiguousarray(g["u_in"]), np.ascontiguousarray(g["rho_in"])
    for label, tune, batched in (("per-step launches, life_step x n", 30, False), ("per-step launches, life_step_n", 30, True),
                                 ("one cluster launch per 1024 steps", 0, True)):
        ctx = capi.Context(capi.Config(tune=tune, **kw))
        ctx.upload_state(f, rho, u, fxy, None, u_in, rho_in)
        ctx.step_n(1, 50)
        ctx.sync()
        n0 = ctx.launch_count()
        t0 = time.perf_counter()
        if batched:
            ctx.step_n(51, steps)
        else:
            for t in range(51, 51 + steps):
                ctx.step(t)
        ctx.sync()
        dt = time.perf_counter() - t0
        vmax, nan, _, _ = ctx.max_speed()
        print("%-16s %4d x %-4d %-40s %7.2f us/step  %8.1f MLUPS  %5d launches  (vmax %.4f%s)"
              % (case, Nx, Ny, label, dt / steps * 1e6, Nx * Ny * steps / dt / 1e6, ctx.launch_count() - n0, vmax, " NaN!" if nan else ""), flush=True)
        ctx.close()
