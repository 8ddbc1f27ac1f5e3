#!/usr/bin/env python3
"""epsilon on the device: time of life_ibm_compute_epsilon (assembly + LU + solve, synchronised) for the UNI_EPSILON systems of the
examples, single-CTA LU (cfg.tune = 40, the round-1 kernel) against the cluster / distributed-shared-memory LU; LAPACK on this host
beside it.  python scripts/eps_timing.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from life_b200 import capi  # noqa: E402
from tests import fixture_state as FS  # noqa: E402

for case in ("TurekHron", "PELskin"):
    g = FS.load(case)
    kw = FS.config_kwargs(g)
    n = len(g["m_ds"])
    res = {}
    for label, tune in (("single CTA, matrix in L2 (round 1)", 40), ("cluster of 8 CTAs, matrix in distributed shared memory", 0)):
        ctx = capi.Context(capi.Config(tune=tune, **kw))
        ctx.ibm_set_markers(g["m_pos"], g["m_vel"], g["m_ds"], np.zeros(n))
        eps = ctx.ibm_compute_epsilon([np.arange(n)])
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(20):
            eps = ctx.ibm_compute_epsilon([np.arange(n)])
        dt = (time.perf_counter() - t0) / 20
        A = ctx.ibm_assemble_epsilon([np.arange(n)])[0]
        err = np.abs(eps - g["m_eps"]).max() / np.abs(g["m_eps"]).max()
        res[label] = eps
        print("%-10s n = %3d  %-56s %8.1f us per call   max rel. difference to the reference's LAPACK epsilon %.1e" % (case, n, label, dt * 1e6, err), flush=True)
        ctx.close()
    import scipy.linalg as sl
    t0 = time.perf_counter()
    for _ in range(20):
        sl.solve(A, np.ones(n))
    print("%-10s n = %3d  %-56s %8.1f us per call" % (case, n, "LAPACK dgesv on this host (scipy, solve only)", (time.perf_counter() - t0) / 20 * 1e6), flush=True)
