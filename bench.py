#!/usr/bin/env python3
"""bench.py — the contract benchmark of life-b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S] [--collision bgk|cm]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): MLUPS of the fused D2Q9 fp64 step = lattice nodes x steps / (1e6 x seconds), the definition
the reference prints (src/Grid.cpp:610).  Workload: BASELINE.json configs[4], the synthetic lid-driven cavity
(WALL_TOP = eVelocity, other walls eWall, omega = 1, lid speed 0.1 lattice units, no bodies, force_xy = 0), S x S nodes
per GPU (S = 16384), slab-decomposed along x: Nx = S*N, Ny = S (weak scaling).  One "step" is one life_step()
= one pass of GridClass::lbmKernel over the whole lattice.

One JSON line on stdout (rank 0):
  value     whole-job MLUPS with the lattice resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e       the same metric through the C ABI with HOST buffers inside the timed region: life_upload_state from pinned
            host arrays (the product of initialiseGrid), K x life_step with the reference main loop's writeInfo scan
            (life_max_speed, scalars back to the host) every tinfo = K/10 steps, and life_download_macro into pinned
            host arrays at the end (what writeVTK reads)
  roofline  dominant kernel (bulk stream+collide sweep): 144 algorithmic bytes per node (9 populations x 8 B read and
            written) x nodes per launch / its average CUDA-event duration inside the timed region, against the measured
            HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference (oracle/_ref/libref_syn_<coll>.so, compiled from /root/reference by
            oracle/Makefile) timed on this box's host cores, rank 0, N = 1 only, on a bounded sample

--impl reference times only that CPU reference (all host threads) and prints the same line shape.

Only the cpu_baseline / --impl reference legs execute anything under oracle/ (as the thing being compared against).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MLUPS (D2Q9, fp64)"
UNIT = "MLUPS"
BYTES_PER_NODE = 144          # 9 populations x 8 B read + 9 x 8 B written
REF_SAMPLE_N = 4096           # the compiled reference's synthetic case is SYN_N x SYN_N (oracle/Makefile SYN_N)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference leg (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_reference(collision, steps, warmup, budget_s):
    """MLUPS of the reference's own GridClass::solver() on a SYN_N^2 sample of the synthetic cavity, all host threads.

    Returns (mlups, ms_per_step, steps_timed, kind, cores, sample).  `steps` = None: as many steps as fit `budget_s`.
    """
    cores = host_cores()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    case = "syn_cm" if collision == "cm" else "syn_bgk"
    from oracle import refharness
    if refharness.available(case):
        ref = refharness.RefCase(case)
        n = ref.Nx * ref.Ny
        kind = "reference"
        sample = ("unmodified reference GridClass::solver() compiled from /root/reference (oracle/_ref/libref_%s.so), "
                  "%dx%d cavity (the reference's 32-bit indices cap it below 15447^2), OpenMP x%d"
                  % (case, ref.Nx, ref.Ny, cores))
        step = ref.step
    else:
        # the compiled reference did not travel: time the C restatement instead (same loops, OpenMP)
        from oracle import oracle as O
        N = REF_SAMPLE_N
        p = O.Params(Nx=N, Ny=N, omega=1.0, wall_top=O.VELOCITY, central_moments=int(collision == "cm"),
                     nu_p=(1.0 / 6.0) / (0.1 * (N - 1)))
        ref = O.Oracle(p)
        n = N * N
        kind = "port"
        sample = "oracle/life_oracle.c restatement, %dx%d cavity, OpenMP x%d" % (N, N, cores)
        step = ref.step
    t0 = time.perf_counter()
    step(max(1, warmup))
    per = (time.perf_counter() - t0) / max(1, warmup)
    if steps is None:
        steps = max(3, min(200, int(budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    step(steps)
    dt = time.perf_counter() - t0
    ref.close()
    return n * steps / dt / 1e6, dt / steps * 1e3, steps, kind, cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded: the whole run (warm-up included) stays within a few minutes whatever K is asked for
    t0 = time.perf_counter()
    mlups, ms, steps, kind, cores, sample = time_reference(args.collision, None if args.steps is None else min(args.steps, 400),
                                                           min(args.warmup, 5), 20.0)
    log("reference arm: %.1f MLUPS, %d steps, %.1f s" % (mlups, steps, time.perf_counter() - t0))
    S = args.size
    line = {
        "impl": "reference", "metric": METRIC, "value": mlups, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 5), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(S, args.gpus, args.collision),
        "cpu_baseline": {"value": mlups, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mlups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(S, n_gpus, collision):
    return {"workload": "synthetic lid-driven cavity %dx%d per GPU (BASELINE.json configs[4]); global lattice %dx%d, "
                        "x-slabs" % (S, S, S * n_gpus, S),
            "Nx": S * n_gpus, "Ny": S, "collision": collision, "omega": 1.0, "lid_speed_lattice": 0.1,
            "walls": "top eVelocity, others eWall", "bodies": 0, "parallelism": "slab%d" % n_gpus,
            "bytes_per_node": BYTES_PER_NODE,
            "l2": "no flush needed: the two population buffers are %.1f GB per GPU, far above the 126 MB L2"
                  % (2 * 9 * 8 * S * S / 1e9)}


# ---------------------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.rows = []
        self.proc = None
        cmd = ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"]
        if uuid:
            cmd += ["-i", uuid]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        inside = [r for (t, r) in self.rows if t_begin <= t <= t_end] or [r for (_, r) in self.rows[-3:]]
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
                power.append(float(p[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def pinned_empty(torch, shape):
    """numpy array over page-locked host memory (cudaHostRegister: torch's pinned allocator rounds sizes up to a power of two)."""
    import numpy as np
    a = np.empty(shape, dtype=np.float64)
    a.fill(0.0)   # touch the pages before locking them
    rc = torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    if int(rc) != 0:
        log("cudaHostRegister failed (%s): host buffer stays pageable" % rc)
    return a


def unpin(torch, a):
    try:
        torch.cuda.cudart().cudaHostUnregister(a.ctypes.data)
    except Exception:
        pass


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from life_b200 import capi, dist as D

    rank, world, local = D.env_ranks()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — liblife_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        nccl_id = D.share_nccl_id()

    S, K, W = args.size, args.steps, args.warmup
    Nx, Ny = S * world, S
    if args.global_nx > 0:          # strong scaling: the global lattice is fixed and cut into `world` slabs
        Nx = args.global_nx
    coll = capi.CENTRAL_MOMENTS if args.collision == "cm" else capi.BGK
    # the reference's scalings for this case (src/Grid.cpp:1257-1260 with height_p = 1, omega = 1, lid 0.1 lattice units)
    Dx = 1.0 / (Ny - 1)
    nu_p = (1.0 / 6.0) / (0.1 * (Ny - 1))
    Dt = (1.0 / np.sqrt(3.0)) ** 2 * Dx * Dx * 0.5 / nu_p
    stream = torch.cuda.Stream(device=dev)
    cfg = capi.Config(Nx=Nx, Ny=Ny, omega=1.0, collision=coll, wall_top=capi.VELOCITY, Dx=Dx, Dt=Dt, Dm=Dx ** 3,
                      device=local, rank=rank, nranks=world, kernel=args.kernel)
    cfg.stream = stream.cuda_stream
    ctx = capi.Context(cfg, nccl_id=nccl_id)
    nxl = ctx.nxl
    nodes_local = nxl * Ny
    nodes_global = Nx * Ny

    # initial state as initialiseGrid leaves it (src/Grid.cpp:999-1058): rho = 1, u = 0, f = f_eq = w.
    # Host image: the whole slab (96 B/node: f in, rho and u out) in pinned memory when this box's RAM allows it for all its
    # ranks; otherwise the slab is streamed through the C ABI's column-range calls from one pinned chunk (same bytes over PCIe).
    w9 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    t0 = time.perf_counter()
    need = nodes_local * 96
    try:
        import psutil
        avail = psutil.virtual_memory().available / max(1, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    except Exception:
        avail = 64e9
    if args.host_chunk_columns > 0:
        hc = min(nxl, args.host_chunk_columns)
    elif need < 0.5 * avail:
        hc = nxl
    else:
        hc = max(1, min(nxl, int(min(0.2 * avail, 4e9) // (Ny * 96))))
    h_f = pinned_empty(torch, (hc, Ny, 9))
    h_f[...] = w9
    h_rho = pinned_empty(torch, (hc, Ny))
    h_u = pinned_empty(torch, (hc, Ny, 2))
    u_in = np.tile(np.array([[0.1, 0.0]]), (Ny, 1))
    host_mode = "whole slab in pinned host memory" if hc == nxl else "streamed in %d-column ranges from one pinned chunk" % hc
    log("[rank %d] host image %.1f GB (%s) built and pinned in %.1f s" % (rank, (h_f.nbytes + h_rho.nbytes + h_u.nbytes) / 1e9,
                                                                       host_mode, time.perf_counter() - t0))

    def upload():
        if hc == nxl:
            ctx.upload_state(h_f, None, None, None, None, u_in, None)
            return
        ctx.upload_begin(u_in, None)
        for il0 in range(0, nxl, hc):
            nc = min(hc, nxl - il0)
            ctx.upload_columns(il0, nc, h_f[:nc])
        ctx.upload_end()

    def download_macro(check=False):
        """rho, u of the whole slab to the host; with `check` returns (mean of rho, all finite) — done outside the timed region"""
        tot, fin = 0.0, True
        for il0 in range(0, nxl, hc):
            nc = min(hc, nxl - il0)
            if hc == nxl:
                ctx.download_macro_into(h_rho, h_u)
            else:
                ctx.download_columns_into(il0, nc, None, h_rho[:nc], h_u[:nc], None)
            if check:
                tot += float(h_rho[:nc].sum())
                fin = fin and bool(np.isfinite(h_rho[:nc]).all())
        return tot / nodes_local, fin

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        return D.max_over_ranks(x, device=dev)

    # ---- device-resident throughput ("value") ------------------------------------------------------------------------
    # nvidia-smi needs a second or more to start on a multi-GPU box: launch it now, use only the samples inside the timed region
    uuid = getattr(torch.cuda.get_device_properties(local), "uuid", None)
    uuid = None if uuid is None else (str(uuid) if str(uuid).startswith("GPU-") else "GPU-" + str(uuid))
    sampler = ClockSampler(uuid) if rank == 0 else None
    upload()
    ctx.step_n(1, W)
    download_macro()                         # allocates the on-demand macroscopic planes outside any timed region
    t_next = W + 1
    barrier()
    launches0 = ctx.launch_count()
    ctx.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tb = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        ctx.step_n(t_next, K)
        e1.record(stream)
    barrier()
    te = time.perf_counter()
    ms_total = reduce_max(e0.elapsed_time(e1))
    clocks = sampler.stop(tb, te) if sampler else None
    bulk_ms, bulk_n = ctx.bulk_kernel_ms()
    ctx.set_profiling(False)
    launches = ctx.launch_count() - launches0
    t_next += K
    vmax, has_nan, _, _ = ctx.max_speed()
    if has_nan or not (0.0 < vmax < 0.2):
        raise SystemExit("bench.py: the lattice is not in a physical state after the timed steps (vmax=%r nan=%r)" % (vmax, has_nan))
    value = nodes_global * K / (ms_total * 1e-3) / 1e6
    bulk_ms = reduce_max(bulk_ms)
    # nodes the profiled launch sweeps: the whole slab at N = 1, the interior (all but the two edge columns) at N > 1
    bulk_nodes = nodes_local if world == 1 else (nxl - 2) * Ny

    # ---- end to end through the C ABI with host buffers ("e2e") ---------------------------------------------------------
    tinfo = max(1, K // 10)
    h2d = nodes_local * 72 + u_in.nbytes
    d2h = nodes_local * 24 + (K // tinfo) * 24
    barrier()
    t0 = time.perf_counter()
    upload()
    t_up = time.perf_counter()
    for t in range(1, K + 1):
        ctx.step(t)
        if t % tinfo == 0:
            vm, nan, _, _ = ctx.max_speed()
            if nan:
                raise SystemExit("bench.py: NaN in the e2e run")
    ctx.sync()
    t_steps = time.perf_counter()
    download_macro()
    barrier()
    t_end = time.perf_counter()
    e2e_s = reduce_max(t_end - t0)
    e2e = nodes_global * K / e2e_s / 1e6
    e2e_parts = {"upload_s": t_up - t0, "steps_and_scans_s": t_steps - t_up, "download_s": t_end - t_steps}
    rho_mean, rho_finite = download_macro(check=True)
    if not rho_finite or abs(rho_mean - 1.0) > 1e-6:
        raise SystemExit("bench.py: downloaded density field is not physical")

    ctx.close()
    for a in (h_f, h_rho, h_u):
        unpin(torch, a)
    del h_f, h_rho, h_u

    # ---- report ------------------------------------------------------------------------------------------------------------
    peaks, peak_src = None, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peaks["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0
    achieved = BYTES_PER_NODE * bulk_nodes / (bulk_ms * 1e-3) / 1e9 if bulk_ms > 0 else None
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%d" % (args.collision, S)
        if key in tr and world == 1:
            traffic = tr[key]["dram_bytes_per_launch"]
    except Exception:
        pass

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        try:
            mlups, ms, steps, kind, cores, sample = time_reference(args.collision, None, 2, 15.0)
            cpu = {"value": mlups, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": sample + ", %d steps, %.0f ms/step" % (steps, ms)}
        except Exception as ex:   # the checker must never take the measurement down
            cpu = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference", "sample": "failed: %r" % (ex,)}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if args.global_nx > 0 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(S, world, args.collision), Nx=Nx, **({"workload": "synthetic lid-driven cavity, global lattice %dx%d cut into %d x-slabs (strong scaling)" % (Nx, Ny, world)} if args.global_nx > 0 else {})),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "kernel": "k_bulk (fused stream+collide sweep)", "kernel_ms": bulk_ms, "launches_timed": bulk_n,
                     "nodes_per_launch": bulk_nodes, "peak_source": peak_src,
                     "step_frac": (BYTES_PER_NODE * nodes_local / (ms_total / K * 1e-3) / 1e9) / peak},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                "seconds": e2e_s, "rank0_breakdown": e2e_parts,
                "host_image": host_mode,
                "what": "life_upload_state(pinned host f) + %d x life_step + life_max_speed every %d steps + "
                        "life_download_macro(pinned host rho,u); per-rank bytes averaged over the steps" % (K, tinfo)},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="life_b200", choices=["life_b200", "reference"])
    ap.add_argument("--size", type=int, default=16384, help="lattice nodes per side per GPU")
    ap.add_argument("--collision", default="bgk", choices=["bgk", "cm"])
    ap.add_argument("--kernel", type=int, default=0, help="LIFE_KERNEL_* (0 = auto)")
    ap.add_argument("--global-nx", type=int, default=0,
                    help="strong scaling: fix the global lattice at GLOBAL_NX x SIZE and cut it into --gpus slabs (default: weak scaling, SIZE x SIZE per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-chunk-columns", type=int, default=0,
                    help="stream the host image through life_upload_columns / life_download_columns in ranges of this many "
                         "columns (0 = whole slab if host RAM allows, else automatic)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.steps is None:
        args.steps = 500
    args.warmup = max(args.warmup, 3)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
