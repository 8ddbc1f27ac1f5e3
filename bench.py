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
  cpu_baseline  the unmodified reference (oracle/_ref/libref_syn_<coll>[_8192].so, compiled from /root/reference by
            oracle/Makefile) timed on this box's host cores by rank 0 at EVERY N (after the other ranks have finished), on bounded
            8192^2 and 4096^2 samples, in a child process with OMP_NUM_THREADS set explicitly; the thread count is the one the
            reference's own counter observed
  parity_check  untimed: small cases cut into the same N slabs, compared with the committed fixtures of the compiled reference

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


def _ref_worker(case, steps, warmup, budget_s, threads):
    """Child process of time_reference: loads ONE compiled-reference case (or the C restatement), times it, prints one JSON line."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import refharness
    collision = "cm" if "_cm" in case else "bgk"
    if refharness.available(case):
        ref = refharness.RefCase(case)
        ref.lib.ref_set_omp_threads(int(threads))         # libgomp may have read an inherited OMP_NUM_THREADS=1 already
        observed = int(ref.lib.ref_omp_threads())         # the reference's own Utils::omp_thread_count (src/Utils.cpp:228-240)
        Nx, Ny, kind = ref.Nx, ref.Ny, "reference"
        what = "unmodified reference GridClass::solver() compiled from /root/reference (oracle/_ref/libref_%s.so)" % case
    else:
        # the compiled reference did not travel: time the C restatement instead (same loops, OpenMP)
        from oracle import oracle as O
        N = 8192 if case.endswith("_8192") else REF_SAMPLE_N
        p = O.Params(Nx=N, Ny=N, omega=1.0, wall_top=O.VELOCITY, central_moments=int(collision == "cm"),
                     nu_p=(1.0 / 6.0) / (0.1 * (N - 1)))
        ref = O.Oracle(p)
        Nx, Ny, kind, observed = N, N, "port", int(threads)
        what = "oracle/life_oracle.c restatement"
    t0 = time.perf_counter()
    ref.step(max(1, warmup))
    per = (time.perf_counter() - t0) / max(1, warmup)
    if steps <= 0:
        steps = max(3, min(200, int(budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    ref.step(steps)
    dt = time.perf_counter() - t0
    ref.close()
    print(json.dumps({"mlups": Nx * Ny * steps / dt / 1e6, "ms_per_step": dt / steps * 1e3, "steps": steps, "kind": kind,
                      "threads": observed, "Nx": Nx, "Ny": Ny, "what": what}), flush=True)
    return 0


def time_reference(collision, steps, warmup, budget_s, size=REF_SAMPLE_N):
    """MLUPS of the reference's own GridClass::solver() on a size^2 sample of the synthetic cavity (4096 or 8192: the cases
    oracle/Makefile compiles), on every host core this process may use.

    Runs in a CHILD process whose OMP_NUM_THREADS is set explicitly: under torch.distributed.run the parent inherits
    OMP_NUM_THREADS=1, and libgomp in a process that has imported torch has read it already.  The child reports the team size
    the reference's own thread counter observed.  Returns a dict (mlups, ms_per_step, steps, kind, threads, Nx, Ny, sample).
    """
    cores = host_cores()
    case = ("syn_cm" if collision == "cm" else "syn_bgk") + ("_8192" if size == 8192 else "")
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OPENBLAS_NUM_THREADS="1")
    for k in ("OMP_PROC_BIND", "OMP_PLACES", "GOMP_CPU_AFFINITY"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--ref-worker", case, str(steps or 0), str(warmup), str(budget_s), str(cores)]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    if p.returncode != 0:
        raise RuntimeError("reference worker failed: " + p.stderr[-500:])
    r = json.loads(p.stdout.strip().splitlines()[-1])
    r["cores"] = cores
    r["sample"] = ("%s, %dx%d lid-driven cavity (the reference's 32-bit indices cap it below 15447^2), OpenMP threads observed %d "
                   "of %d host cores, %d steps, %.0f ms/step" % (r["what"], r["Nx"], r["Ny"], r["threads"], cores, r["steps"], r["ms_per_step"]))
    return r


def cpu_baseline_block(collision, budget_s=12.0):
    """cpu_baseline of the JSON line: the 8192^2 sample is the value, the 4096^2 sample is listed beside it."""
    out = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "reference", "sample": None, "samples": []}
    try:
        for size, budget in ((8192, budget_s), (REF_SAMPLE_N, budget_s / 2)):
            r = time_reference(collision, None, 1, budget, size)
            out["samples"].append({"lattice": "%dx%d" % (r["Nx"], r["Ny"]), "value": r["mlups"], "steps": r["steps"],
                                   "ms_per_step": r["ms_per_step"], "threads": r["threads"], "kind": r["kind"]})
            if out["value"] is None:
                out.update(value=r["mlups"], kind=r["kind"], sample=r["sample"], cores=r["threads"])
    except Exception as ex:   # the checker must never take the measurement down
        out["sample"] = "failed: %r" % (ex,)
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded: the whole run (warm-up included) stays within a few minutes whatever K is asked for
    t0 = time.perf_counter()
    W = min(args.warmup, 3)
    K = None if args.steps is None else min(args.steps, 60)
    big = time_reference(args.collision, K, W, 20.0, 8192)
    small = time_reference(args.collision, None if K is None else min(4 * K, 200), W, 8.0, REF_SAMPLE_N)
    log("reference arm: %.1f MLUPS at 8192^2 (%d steps), %.1f MLUPS at 4096^2, %d OpenMP threads, %.1f s"
        % (big["mlups"], big["steps"], small["mlups"], big["threads"], time.perf_counter() - t0))
    cfg = workload_config(args.size, args.gpus, args.collision)
    cfg["workload"] = ("CPU SAMPLE of BASELINE.json configs[4]: the synthetic lid-driven cavity at %dx%d (not %dx%d: the reference's "
                       "32-bit indices overflow at 15447^2 and it needs 228 B/node of host RAM), same walls / omega / lid speed; MLUPS is "
                       "size-normalised" % (big["Nx"], big["Ny"], cfg["Nx"], cfg["Ny"]))
    cfg.update(Nx=big["Nx"], Ny=big["Ny"], parallelism="OpenMP x%d, one process" % big["threads"], gpu_arm_lattice="%dx%d" % (args.size * args.gpus, args.size))
    cfg.pop("l2", None)
    samples = [{"lattice": "%dx%d" % (r["Nx"], r["Ny"]), "value": r["mlups"], "steps": r["steps"], "ms_per_step": r["ms_per_step"],
                "threads": r["threads"], "kind": r["kind"]} for r in (big, small)]
    line = {
        "impl": "reference", "metric": METRIC, "value": big["mlups"], "unit": UNIT, "n_gpus": args.gpus, "steps": big["steps"],
        "warmup": W, "ms_per_step": big["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": big["mlups"], "unit": UNIT, "cores": big["threads"], "kind": big["kind"], "sample": big["sample"],
                         "samples": samples},
        "e2e": {"value": big["mlups"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def parity_spot_check(capi, D, dist, rank, world, local, dev):
    """Untimed slab-parity check on the SAME ranks / communicator layout as the benchmark (the driver's test box has one GPU, so
    the multi-GPU path would otherwise only be exercised by the timing): small cases cut into `world` slabs, every rank's slab
    compared with the COMMITTED FIXTURES the compiled reference wrote (tests/golden/*.npz: sampled fields, marker forces of the
    recorded FSI trace) — no oracle involved.
      t_periodic_cm   fast kernels, central moments, periodic ring        rho, u, f at the sampled nodes <= 1e-10
      t_periodic_bgk, t_womersley, t_convective   cfg.exact              the reference's doubles, bit for bit
      Cylinder, Honami  (bodies; supports straddle the slab faces)       trace forces and sampled fields <= 1e-10
    Returns the dict that goes into the JSON line as "parity_check"."""
    import numpy as np
    import torch
    from tests import fixture_state as FS

    def rel(a, b, floor=0.0):
        a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
        n = max(np.linalg.norm(b), floor * np.sqrt(max(b.size, 1)))
        return float(np.linalg.norm(a - b) / n) if n > 0 else float(np.linalg.norm(a - b))

    out, ok_all = {}, True
    plan = [("t_periodic_cm", 0), ("t_periodic_bgk", 1), ("t_womersley", 1), ("t_convective", 1), ("Cylinder", 0), ("Honami", 0)]
    for case, exact in plan:
        g = FS.load(case)
        Nx = int(g["Nx"])
        if Nx // world < 4:
            out[case] = "skipped (slabs thinner than 4 columns)"
            continue
        f, rho, u, fxy, u_in, rho_in = FS.initial_state(g)
        cfg = capi.Config(device=local, rank=rank, nranks=world, exact=exact, **FS.config_kwargs(g))
        ctx = capi.Context(cfg, nccl_id=D.share_nccl_id() if world > 1 else None)
        b, e = ctx.i_begin, ctx.i_end
        ctx.upload_state(f[b:e], rho[b:e], u[b:e], fxy[b:e], None, u_in, rho_in)
        worst = 0.0
        steps = int(g["steps"])
        if "trace_step" in g.files:
            k = 0
            for t in range(1, steps + 1):
                ctx.step(t)
                while True:
                    ctx.ibm_set_markers(g["trace_pos"][k], g["trace_vel"][k], g["trace_ds"][k], g["trace_eps"][k])
                    worst = max(worst, rel(ctx.ibm_interp(), g["trace_force"][k], 1e-6))
                    k += 1
                    if g["trace_last"][k - 1]:
                        break
                ctx.ibm_spread()
        else:
            ctx.step_n(1, steps)
        st = ctx.download_state()
        ctx.close()
        m, il, j = FS.sample_index(g, b, e)
        bitwise = True
        for name in ("rho", "u", "f") + (("force_ibm",) if "trace_step" in g.files else ()):
            if m.any():
                got, want = st[name][il, j], g[name][m]
                worst = max(worst, rel(got, want, 1e-12 if name == "force_ibm" else (1e-5 if name == "u" else 0.0)))
                bitwise = bitwise and bool(np.array_equal(got, want))
        ok = (bitwise if exact else worst < 1e-10)
        flag = torch.tensor([0.0 if ok else 1.0, worst], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        ok, worst = flag[0].item() == 0.0, float(flag[1].item())
        out[case] = {"slabs": world, "steps": steps, "mode": "exact: bitwise" if exact else "fast: rel L2 <= 1e-10",
                     "worst_rel_l2": worst, "ok": ok}
        ok_all = ok_all and ok
    out["ok"] = ok_all
    out["against"] = "tests/golden/*.npz written by the compiled reference (oracle/_ref), sampled nodes of every rank's slab + FSI trace forces"
    return out


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
                      device=local, rank=rank, nranks=world, kernel=args.kernel, inplace=1 if args.inplace else 0, exact=1 if args.exact else 0)
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

    # ---- untimed: slab parity on this very rank layout ---------------------------------------------------------------------
    parity = None
    if not args.no_parity_check:
        try:
            parity = parity_spot_check(capi, D, dist, rank, world, local, dev)
        except Exception as ex:
            parity = {"ok": False, "error": repr(ex)}
        if rank == 0:
            log("parity spot check on %d slab(s): %s" % (world, "ok" if parity.get("ok") else "FAILED %r" % (parity,)))

    # ---- report ------------------------------------------------------------------------------------------------------------
    peaks, peak_src = None, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peaks["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0
    achieved = BYTES_PER_NODE * bulk_nodes / (bulk_ms * 1e-3) / 1e9 if bulk_ms > 0 else None
    traffic, traffic_src = None, "none for this configuration"
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%d%s" % (args.collision, S, "_inplace" if args.inplace else "")
        if key in tr and world == 1:
            traffic = tr[key]["dram_bytes_per_launch"]
            traffic_src = "static: ncu --set full capture of this kernel on this workload, %s (profiles/traffic.json); not measured in this run" % tr[key].get("source", "profiles/")
    except Exception:
        pass

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    # the reference's OpenMP build on this box's host cores, beside the GPU number at EVERY N (rank 0 alone, after the other
    # ranks have finished: they exit while this runs, so the sample has the host to itself)
    torch.cuda.empty_cache()
    if world > 1 and not args.no_cpu_baseline:
        time.sleep(3.0)      # the other ranks are unmapping their 26 GB host images while they exit: keep that out of the CPU sample
    cpu = None if args.no_cpu_baseline else cpu_baseline_block(args.collision)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if args.global_nx > 0 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(S, world, args.collision), Nx=Nx, layout="in place, one population buffer (cfg.inplace)" if args.inplace else "two population buffers",
                       arithmetic="reference operation order, no FMA contraction (cfg.exact)" if args.exact else "factored collision, FMA", **({"workload": "synthetic lid-driven cavity, global lattice %dx%d cut into %d x-slabs (strong scaling)" % (Nx, Ny, world)} if args.global_nx > 0 else {})),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
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
        "parity_check": parity,
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
    ap.add_argument("--inplace", action="store_true", help="cfg.inplace: one population buffer (72 B/node resident), in-place shift sweep")
    ap.add_argument("--exact", action="store_true", help="cfg.exact: the step in the reference's operation order (bitwise its results)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--ref-worker", nargs=5, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--host-chunk-columns", type=int, default=0,
                    help="stream the host image through life_upload_columns / life_download_columns in ranges of this many "
                         "columns (0 = whole slab if host RAM allows, else automatic)")
    args = ap.parse_args()
    if args.ref_worker:
        case, steps, warmup, budget, threads = args.ref_worker
        return _ref_worker(case, int(steps), int(warmup), float(budget), int(threads))
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.steps is None:
        args.steps = 500
    args.warmup = max(args.warmup, 3)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
