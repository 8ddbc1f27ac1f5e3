#!/usr/bin/env python3
"""Where the start-up time of a small case goes: CUDA context, library load, life_create, first upload, first steps."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
rt = C.CDLL("libcudart.so.12")
t1 = time.perf_counter()
rt.cudaFree(None)                     # creates the primary context
t2 = time.perf_counter()
import numpy as np  # noqa: E402
from life_b200 import capi  # noqa: E402
capi.load()
t3 = time.perf_counter()
N = 101
cfg = capi.Config(Nx=N, Ny=N, omega=1.9, collision=capi.CENTRAL_MOMENTS, wall_top=capi.VELOCITY, Dx=0.01, Dt=0.001, Dm=1e-6)
ctx = capi.Context(cfg)
t4 = time.perf_counter()
f = np.empty((N, N, 9)); f[...] = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
ctx.upload_state(f, np.ones((N, N)), np.zeros((N, N, 2)), None, None, np.tile(np.array([[0.1, 0.0]]), (N, 1)), None)
t5 = time.perf_counter()
ctx.step(1); ctx.sync()
t6 = time.perf_counter()
ctx.step_n(2, 499); ctx.sync()
t7 = time.perf_counter()
ctx.max_speed()
t8 = time.perf_counter()
print("dlopen libcudart %.3f s | cudaFree(0) = driver init + primary context %.3f s | load liblife_b200.so %.3f s | life_create %.3f s | "
      "first upload %.3f s | first step (module load of the kernels it uses) %.3f s | 499 steps %.4f s | first max_speed %.4f s"
      % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7))
