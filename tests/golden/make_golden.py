#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the compiled, UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py Cylinder   # one case

The reference ships no golden data (its own protocol, testing/store-ref-data.sh, generates RefData from
the reference build itself); this script does the same, in process, through oracle/ref_harness.cpp.

One .npz per case:
  case description  : Nx, Ny, walls, flags and every params.h value the hot path reads (enough to configure the oracle
                      and the CUDA library without the reference)
  init_*            : the state initialiseGrid produced, SAMPLED (see `sample`)
  steps             : number of time steps run
  sample            : node ids (i*Ny + j) at which fields are stored — a stride sample plus every boundary-adjacent
                      corner region and every IBM support site, capped so a fixture stays a few hundred KB
  rho, u, f, force_ibm at `sample` after `steps` steps; plus whole-field sums (a checksum over every node)
  type_bc           : BCVec (exact integer map)
  markers           : pos, vel, ds, epsilon, body, flex at t = 0; supports at t = 0 (exact integer map + delta)
  trace_*           : for cases with bodies, per sub-iteration of every step: the marker pos / vel / ds / epsilon the host
                      handed to ibmKernelInterp and the marker force it got back (the FSI call-site trace that lets the
                      kernels be checked without the host FEM solver), and which sub-iteration closed each step
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refharness import RefCase, available  # noqa: E402
from tests.initstate import wavy_state  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# case -> (steps, max samples)
CASES = {
    "LidDrivenCavity": (200, 2500),
    "ChannelFlow": (200, 2500),
    "Cylinder": (100, 2500),
    "TurekHron": (40, 2500),
    "InvertedFlag": (30, 2500),
    "Honami": (2, 2500),
    "PELskin": (30, 2500),
    "t_periodic_bgk": (100, 1500),
    "t_periodic_cm": (100, 1500),
    "t_convective": (150, 1500),
    "t_freeslip_cm": (150, 1500),
    "t_womersley": (150, 1500),
    "t_velocity_box": (100, 1500),
    "t_pressure_left": (150, 1500),
    "t_yperiodic": (100, 1500),
}


WAVY = {"Cylinder": (0.02, 0.01), "TurekHron": (0.02, 0.01), "Honami": (0.02, 0.01)}


def pick_samples(ref, cap):
    Nx, Ny = ref.Nx, ref.Ny
    N = Nx * Ny
    ids = set()
    # corners and their neighbourhood (where the boundary logic is most intricate)
    for i0 in (0, Nx - 4):
        for j0 in (0, Ny - 4):
            for di in range(4):
                for dj in range(4):
                    ids.add((i0 + di) * Ny + (j0 + dj))
    # support sites of the markers at t = 0
    if ref.n_markers:
        cnt, idx, jdx, _ = ref.supports()
        for m in range(0, ref.n_markers, max(1, ref.n_markers // 60)):
            for s in range(cnt[m]):
                ids.add(int(idx[m, s]) * Ny + int(jdx[m, s]))
    # mid-edge boundary nodes and their inward neighbours
    for j in (0, 1, 2, Ny - 3, Ny - 2, Ny - 1):
        for i in range(0, Nx, max(1, Nx // 40)):
            ids.add(i * Ny + j)
    for i in (0, 1, 2, Nx - 3, Nx - 2, Nx - 1):
        for j in range(0, Ny, max(1, Ny // 40)):
            ids.add(i * Ny + j)
    rest = cap - len(ids)
    if rest > 0:
        stride = max(1, N // rest)
        ids.update(range(stride // 2, N, stride))
    return np.array(sorted(ids), dtype=np.int64)


def describe(ref):
    return dict(Nx=ref.Nx, Ny=ref.Ny, walls=np.array(ref.walls, np.int32), flags=ref.flags,
                central_moments=int(ref.central_moments), ordered=int(ref.ordered),
                uni_epsilon=int(bool(ref.flags & ref.FLAG_UNI_EPS)), profile=ref.profile,
                omega=ref.omega, Dx=ref.Dx, Dt=ref.Dt, Dm=ref.Dm, Drho=ref.Drho, inlet_ramp=ref.inlet_ramp,
                womersley=ref.womersley, dpdx=ref.dpdx, dpdy=ref.dpdy, gravityX=ref.gravityX, gravityY=ref.gravityY,
                height_p=ref.height_p, nu_p=ref.nu_p, rho_p=ref.rho_p, subTol=ref.subTol, uxInlet_p=ref.uxInlet_p,
                uyInlet_p=ref.uyInlet_p, ux0_p=ref.ux0_p, uy0_p=ref.uy0_p, has_ibm=int(ref.has_ibm),
                has_flex=int(ref.has_flex))


def make(case):
    steps, cap = CASES[case]
    ref = RefCase(case)
    d = describe(ref)
    # The reference can only start from a uniform / profile state (initialiseGrid).  The extra cases overwrite it with a
    # smooth non-uniform, non-equilibrium state so that streaming, wrap-around and every boundary type see real
    # gradients.  The three examples that start from rest behind a long inlet ramp (Cylinder, TurekHron, Honami) get a
    # gentler version of the same: from rest their velocities and marker forces stay below 1e-11 for hundreds of steps,
    # i.e. below the rounding of the O(1) populations, where a relative comparison means nothing.
    amp, neq = WAVY.get(case, (0.04, 0.02) if case.startswith("t_") else (0.0, 0.0))
    d["wavy"] = int(amp > 0.0)
    d["wavy_amp"], d["wavy_neq"] = amp, neq
    if d["wavy"]:
        f0, rho0, u0 = wavy_state(ref.Nx, ref.Ny, ref.central_moments, amp=amp, non_equilibrium=neq)
        ref.set_state(f=f0, rho=rho0, u=u0)
    sample = pick_samples(ref, cap)
    Ny = ref.Ny
    si, sj = sample // Ny, sample % Ny
    d.update(steps=steps, sample=sample)
    d["init_f"] = ref.f()[si, sj]
    d["init_u"] = ref.u()[si, sj]
    d["init_rho"] = ref.rho()[si, sj]
    d["init_force_xy0"] = ref.force_xy()[0, 0]
    d["u_in"] = ref.u_in()
    d["rho_in"] = ref.rho_in()
    d["bcvec"] = ref.bcvec().astype(np.int64)
    d["type_edges"] = np.concatenate([ref.type()[0], ref.type()[-1], ref.type()[:, 0], ref.type()[:, -1]]).astype(np.int8)
    if ref.n_markers:
        m = ref.markers()
        cnt, idx, jdx, dirac = ref.supports()
        d.update(m_pos=m["pos"], m_vel=m["vel"], m_ds=m["ds"], m_eps=m["epsilon"], m_body=m["body"], m_flex=m["flex"],
                 s_count=cnt, s_idx=idx, s_jdx=jdx, s_dirac=dirac)

    tr_pos, tr_vel, tr_ds, tr_eps, tr_force, tr_step, tr_last = [], [], [], [], [], [], []
    for _ in range(steps):
        ref.t = ref.t + 1
        ref.lbm_kernel()
        if ref.has_ibm:
            # ObjectsClass::objectKernel (Objects.cpp:26-60), stage by stage, recording what crosses the seam
            ref.subit = 0
            while True:
                if ref.has_flex:
                    ref.recompute_object_vals()
                mk = ref.markers()
                ref.ibm_interp()
                force = ref.markers()["force"]
                tr_pos.append(mk["pos"]); tr_vel.append(mk["vel"]); tr_ds.append(mk["ds"]); tr_eps.append(mk["epsilon"])
                tr_force.append(force); tr_step.append(ref.t)
                if not ref.has_flex:
                    tr_last.append(1)
                    break
                ref.fem_kernel()
                ref.subit = ref.subit + 1
                done = not (ref.subit < 20 and ref.subres > ref.subTol)
                tr_last.append(int(done))
                if done:
                    break
            ref.ibm_spread()
    d["t_end"] = ref.t
    d["rho"] = ref.rho()[si, sj]
    d["u"] = ref.u()[si, sj]
    f = ref.f()
    d["f"] = f[si, sj]
    d["force_ibm"] = ref.force_ibm()[si, sj]
    # whole-field checksums: cover every node, not just the sample
    d["sum_rho"] = float(ref.rho().sum())
    d["sum_f_per_v"] = f.reshape(-1, 9).sum(axis=0)
    d["sum_abs_u"] = float(np.abs(ref.u()).sum())
    d["sum_u2"] = float((ref.u() ** 2).sum())
    if ref.has_ibm:
        d.update(trace_pos=np.array(tr_pos), trace_vel=np.array(tr_vel), trace_ds=np.array(tr_ds),
                 trace_eps=np.array(tr_eps), trace_force=np.array(tr_force), trace_step=np.array(tr_step, np.int32),
                 trace_last=np.array(tr_last, np.int8), final_force=ref.markers()["force"])
    ref.close()
    path = os.path.join(OUT, case + ".npz")
    np.savez_compressed(path, **d)
    print("%-18s steps=%-4d samples=%-5d subits=%-4d %7.1f KB" % (case, steps, len(sample), len(tr_step),
                                                                os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for c in names:
        if not available(c):
            print("skip", c, "(oracle/_ref/libref_%s.so not built)" % c)
            continue
        # one process per case: the harness holds one compile-time case per shared object, and the reference calls
        # exit(99) on its own errors
        if len(names) > 1:
            import subprocess
            subprocess.check_call([sys.executable, os.path.abspath(__file__), c])
        else:
            make(c)
