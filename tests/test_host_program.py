"""The drop-in at program level: LIFE's own main() with the hot path on the B200 (life_b200/host/_build/<case>/LIFE_b200)
against the unmodified reference program (LIFE_ref, same sources, same case) — the reference's own regression protocol
(testing/store-ref-data.sh, testing/run-tests.sh: 500 steps, restart every 100, TurekHron run twice to exercise the restart
path, then compare Results/), with `diff -r` relaxed to the north-star tolerance: fields and marker forces within
relative L2 1e-10.  The host-side FEM / LAPACK epsilon solve / Aitken loop run live in both programs.
"""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from tests import cases as K
from tests import restartfile as R
from tests import vtkfile as V

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "life_b200", "host", "_build")
EXAMPLES = K.EXAMPLES_LBM + K.EXAMPLES_IBM


def _have(case, exe):
    return os.path.exists(os.path.join(HOST, case, exe))


def _run(case, exe, workdir, times=1, **extra_env):
    os.makedirs(workdir, exist_ok=True)
    inp = os.path.join(HOST, case, "input")
    if os.path.isdir(inp):
        shutil.copytree(inp, os.path.join(workdir, "input"), dirs_exist_ok=True)
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", **extra_env)
    for _ in range(times):
        p = subprocess.run([os.path.join(HOST, case, exe)], cwd=workdir, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=1200)
    m = re.search(r"Simulation took ([0-9.]+) seconds", p.stdout)          # main.cpp:93, includes construction and all I/O
    p.seconds = float(m.group(1)) if m else float("nan")
    return p


FLEXIBLE = ["TurekHron", "InvertedFlag", "Honami", "PELskin"]     # live FEM + Aitken sub-iterations between interp and spread


def _compare(case, ref_dir, new_dir):
    """Relative L2 differences between two Results/ trees: fields (Fluid.restart), markers (IBM.restart), TotalForces.out."""
    a = R.read_fluid(os.path.join(ref_dir, "Results", "Restart", "Fluid.restart"))
    b = R.read_fluid(os.path.join(new_dir, "Results", "Restart", "Fluid.restart"))
    assert (a["t"], a["Nx"], a["Ny"]) == (b["t"], b["Nx"], b["Ny"])
    err = {name: float(K.rel_l2(b[name], a[name])) for name in ("rho", "u", "f")}
    err["force_ibm"] = float(K.rel_l2(b["force_ibm"], a["force_ibm"], floor=1e-12))
    ibm = os.path.join(ref_dir, "Results", "Restart", "IBM.restart")
    if os.path.exists(ibm):
        ma, mb = R.read_ibm(ibm), R.read_ibm(os.path.join(new_dir, "Results", "Restart", "IBM.restart"))
        assert [x["id"] for x in ma] == [x["id"] for x in mb]
        cat = lambda m, k: np.concatenate([x[k] for x in m])
        err["marker_pos"] = float(K.rel_l2(cat(mb, "pos"), cat(ma, "pos")))
        err["marker_vel"] = float(K.rel_l2(cat(mb, "vel"), cat(ma, "vel"), floor=1e-7))   # a body at rest: rounding noise
        err["marker_force"] = float(K.rel_l2(cat(mb, "force"), cat(ma, "force"), floor=1e-9))
        ta = R.read_table(os.path.join(ref_dir, "Results", "TotalForces.out"))
        tb = R.read_table(os.path.join(new_dir, "Results", "TotalForces.out"))
        assert ta.shape == tb.shape and np.array_equal(ta[:, 0], tb[:, 0])
        err["TotalForces.out"] = float(K.rel_l2(tb[:, 2:4], ta[:, 2:4]))
    # the last fluid VTK file (written from the device state by life_write_vtk, or by the reference's writer with LIFE_B200_HOST_IO):
    # same size, same XML head / tail, fields within tolerance
    va = os.path.join(ref_dir, "Results", "VTK", "Fluid.%d.vti" % a["t"])
    vb = os.path.join(new_dir, "Results", "VTK", "Fluid.%d.vti" % a["t"])
    if os.path.exists(va):
        fa, fb = V.read_fluid(va, a["Nx"], a["Ny"]), V.read_fluid(vb, a["Nx"], a["Ny"])
        assert os.path.getsize(va) == os.path.getsize(vb)
        ra, rb = open(va, "rb").read(), open(vb, "rb").read()
        cut = ra.index(b"_") + 1
        assert ra[:cut] == rb[:cut] and ra[-40:] == rb[-40:]
        err["vti density"] = float(K.rel_l2(fb["density"], fa["density"]))
        # (pressure is an affine map of density that subtracts the mean: its relative error is the density's times |rho| / |rho - rho_0|;
        #  tests/test_gpu_output.py holds all three blocks to the reference's bytes)
        err["vti velocity"] = float(K.rel_l2(fb["velocity"], fa["velocity"]))
    return a["t"], err


@pytest.mark.gpu
@pytest.mark.parametrize("device_eps", [0, 1, 2, 3, "host-io", "device-fem", "device-fem-resident"],
                         ids=["host-eps", "device-assembly+lapack", "device-eps", "auto-eps", "host-io", "device-fem", "device-fem-resident"])
@pytest.mark.parametrize("case", EXAMPLES)
def test_program_reproduces_reference_results(case, device_eps, tmp_path):
    """Default build of the drop-in: device-fed files (life_write_vtk / life_write_restart / life_read_restart / life_max_speed);
    the "host-io" variant (LIFE_B200_HOST_IO=1) downloads into the host mirrors and runs the reference's own writers."""
    if not (_have(case, "LIFE_b200") and _have(case, "LIFE_ref")):
        pytest.skip("life_b200/host/_build/%s not built (make -C life_b200/host needs /root/reference)" % case)
    host_io = device_eps == "host-io"
    if host_io:
        device_eps = 0
        if case not in ("ChannelFlow", "TurekHron"):
            pytest.skip("the host-mirror I/O variant is exercised on one plain and one restarted body case")
    resident = device_eps == "device-fem-resident"     # LIFE_B200_DEVICE_FEM=2: the whole sub-iteration loop on the device
    device_fem = device_eps in ("device-fem", "device-fem-resident")
    if device_fem:
        device_eps = 1
        if case not in FLEXIBLE:
            pytest.skip("the structural solver only runs for flexible bodies")
    if device_eps and case not in FLEXIBLE:
        pytest.skip("epsilon is only recomputed for flexible bodies")
    times = 2 if case == "TurekHron" else 1          # second run restarts from Results/Restart (store-ref-data.sh:51-53)
    ref = _run(case, "LIFE_ref", str(tmp_path / "ref"), times)
    assert ref.returncode == 0, ref.stdout[-2000:]
    exe = "LIFE_b200"
    new = _run(case, exe, str(tmp_path / "b200"), times, LIFE_B200_DEVICE_EPSILON=str(device_eps),
               LIFE_B200_HOST_IO="1" if host_io else "0", LIFE_B200_DEVICE_FEM=("2" if resident else "1") if device_fem else "0")
    assert new.returncode == 0, new.stdout[-2000:] + new.stderr[-2000:]
    assert ("life_fem_dynamic calls" in new.stderr) == device_fem
    assert "life_step" in new.stderr and " 0 life_step" not in new.stderr      # the CUDA path really ran
    assert (" 0 life_ibm_compute_epsilon" in new.stderr) == (not device_eps)
    if resident:
        assert " 0 life_ibm_interp" not in new.stderr and "us per sub-iteration" in new.stderr
    print("\n%s wall time of the last run (500 steps incl. all host work and I/O): LIFE_ref %.2f s (%d host threads), LIFE_b200 %.2f s   %s"
          % (case, ref.seconds, os.cpu_count(), new.seconds, new.stderr.strip().splitlines()[-1]))
    t_end, err = _compare(case, str(tmp_path / "ref"), str(tmp_path / "b200"))
    assert t_end == 500 * times
    # the report of writeInfo (src/Grid.cpp:590-616; from life_max_speed in the drop-in): same lines, same numbers as printed
    info = lambda out, pat: re.findall(pat, out)
    for pat in (r"Time step (\d+) of (\d+)", r"Simulation has done (\S+) of (\S+) seconds"):
        assert info(new.stdout, pat) == info(ref.stdout, pat) and len(info(ref.stdout, pat)) == 51, pat
    for pat in (r"Max Velocity = (\S+)", r"Max Velocity \(m/s\) = (\S+)", r"Max Reynolds number = (\S+)"):
        a, b = np.array(info(ref.stdout, pat), float), np.array(info(new.stdout, pat), float)
        assert a.shape == b.shape == (51,), pat
        if case not in FLEXIBLE:     # (flexible bodies: see below)
            assert np.allclose(b, a, rtol=2e-4, atol=1e-12), (pat, a, b)
    first = [str(tmp_path / d / "Results" / "VTK" / "Fluid.0.vti") for d in ("ref", "b200")]
    if all(os.path.exists(x) for x in first):          # the initial state: identical bytes
        assert open(first[0], "rb").read() == open(first[1], "rb").read()
    assert sorted(os.listdir(tmp_path / "ref" / "Results" / "VTK")) == sorted(os.listdir(tmp_path / "b200" / "Results" / "VTK"))

    print("\n%s LIFE_b200 vs LIFE_ref: %s" % (case, err))
    if case in FLEXIBLE:
        # With flexible bodies the host's Aitken-relaxed sub-iteration loop (src/Objects.cpp:33-52, converged only to subTol =
        # 1e-4 .. 1e-8) sits between interp and spread, and the coupled system amplifies ANY rounding-level perturbation
        # exponentially: the unmodified reference recompiled with -mfma differs from itself by 2e-8 (Honami), 7e-9 (PELskin),
        # 9e-12 (TurekHron) in TotalForces.out after 50 steps and by 0.4 / 5e-5 / 5e-7 after 500 (measured on the CPU,
        # DESIGN.md §2).  A rounding-equivalent kernel (the default factored collision, the device LU / FEM variants) therefore
        # has no defined 500-step tolerance; what is defined, and asserted, is (a) the early trajectory — every TotalForces.out
        # row up to t = 20 within 1e-7 of the reference (10 printed digits, params.h:110) — and (b) BITWISE equality of the whole
        # run in exact mode, test_program_in_exact_mode_is_bitwise_the_reference below.  The end-of-run differences are printed.
        ta = R.read_table(str(tmp_path / "ref" / "Results" / "TotalForces.out"))
        tb = R.read_table(str(tmp_path / "b200" / "Results" / "TotalForces.out"))
        early = (ta[:, 0] > 0) & (ta[:, 0] <= 20)
        assert early.sum() == 2
        scale = np.abs(ta[:, 1:]).max()
        assert np.abs(tb[early, 1:] - ta[early, 1:]).max() <= 1e-7 * scale, (case, ta[early], tb[early])
        return
    # TotalForces.out is printed with 10 significant digits (params.h:110)
    bar = {k: (1e-8 if k == "TotalForces.out" else K.TOL) for k in err}
    bad = {k: (err[k], bar[k]) for k in err if not err[k] <= bar[k]}
    assert not bad, (case, bad)


def _diff_r(ref_dir, new_dir):
    """testing/run-tests.sh:100 — `diff -r Results RefData/Results --exclude=Log.out`: the list of files that differ."""
    a_root, b_root = os.path.join(ref_dir, "Results"), os.path.join(new_dir, "Results")
    listing = lambda root: sorted(os.path.relpath(os.path.join(d, f), root) for d, _, fs in os.walk(root) for f in fs if f != "Log.out")
    la, lb = listing(a_root), listing(b_root)
    assert la == lb, (sorted(set(la) ^ set(lb)))
    return [f for f in la if open(os.path.join(a_root, f), "rb").read() != open(os.path.join(b_root, f), "rb").read()], len(la)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["default", "host-eps", "host-io", "inplace"])
@pytest.mark.parametrize("case", EXAMPLES)
def test_program_in_exact_mode_is_bitwise_the_reference(case, variant, tmp_path):
    """The reference's own regression protocol, unrelaxed: LIFE_b200 with LIFE_B200_EXACT=1 (cfg.exact: the step in the
    reference's operation order, kernels of namespace life::exact) against LIFE_ref — 500 steps, TurekHron twice (restart), then
    `diff -r Results` excluding Log.out must find NO differing file: every .vti / .vtp, Fluid / IBM / FEM.restart,
    TotalForces.out, TipPositions.out ... byte for byte, with the live FEM + Aitken sub-iteration loop in between
    (TurekHron, InvertedFlag, Honami, PELskin).  Six cases are BGK, LidDrivenCavity is central moments (the reference's nine
    expanded polynomials restated term by term, d2q9.cuh: collide_cm_ref): all seven examples of the reference."""
    if not (_have(case, "LIFE_b200") and _have(case, "LIFE_ref")):
        pytest.skip("life_b200/host/_build/%s not built (make -C life_b200/host needs /root/reference)" % case)
    if variant not in ("default", "inplace") and case not in ("ChannelFlow", "TurekHron", "PELskin"):
        pytest.skip("the epsilon / host-mirror variants are exercised on one plain case and the two UNI_EPSILON restart / Womersley cases")
    times = 2 if case == "TurekHron" else 1
    ref = _run(case, "LIFE_ref", str(tmp_path / "ref"), times)
    assert ref.returncode == 0, ref.stdout[-2000:]
    # default = epsilon matrix assembled on the device + the host's LAPACK, device-fed files; host-eps = the reference's own
    # computeEpsilon; host-io = downloads into the host mirrors + the reference's own writers / reader
    new = _run(case, "LIFE_b200", str(tmp_path / "b200"), times, LIFE_B200_EXACT="1",
               LIFE_B200_DEVICE_EPSILON="0" if variant == "host-eps" else "1", LIFE_B200_HOST_IO="1" if variant == "host-io" else "0",
               LIFE_B200_INPLACE="1" if variant == "inplace" else "0")      # inplace: one population buffer, shifted layout (cfg.inplace)
    assert new.returncode == 0, new.stdout[-2000:] + new.stderr[-2000:]
    assert "life_step" in new.stderr and " 0 life_step" not in new.stderr      # the CUDA path really ran
    differing, n_files = _diff_r(str(tmp_path / "ref"), str(tmp_path / "b200"))
    t_end, err = _compare(case, str(tmp_path / "ref"), str(tmp_path / "b200"))
    print("\n%s exact mode: %d files compared, differing: %s; field differences %s" % (case, n_files, differing, err))
    assert t_end == 500 * times
    assert not differing, (case, differing, err)
    a = R.read_fluid(str(tmp_path / "ref" / "Results" / "Restart" / "Fluid.restart"))
    b = R.read_fluid(str(tmp_path / "b200" / "Results" / "Restart" / "Fluid.restart"))
    for name in ("rho", "u", "f", "force_ibm"):
        assert np.array_equal(a[name], b[name]), (case, name)
    # the numbers writeInfo prints (max velocity from life_max_speed) are the same text
    for pat in (r"Max Velocity = (\S+)", r"Max Velocity \(m/s\) = (\S+)", r"Max Reynolds number = (\S+)"):
        assert re.findall(pat, new.stdout) == re.findall(pat, ref.stdout), pat


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4])
@pytest.mark.parametrize("case", ["ChannelFlow", "Cylinder", "PELskin", "Honami"])
def test_program_on_several_gpus(case, gpus, tmp_path):
    """LIFE's own main() and host code around N slabs on N GPUs (LIFE_B200_GPUS=N: one host thread per rank inside life_host.cpp,
    NCCL halo exchange, every rank writing its byte ranges of the shared fluid files) — multi-GPU without Python.
    Body-free case, exact mode: bit for bit the reference program (per-node arithmetic does not depend on the decomposition).
    Bodies: a marker whose support straddles a slab face is gathered as two partial sums added by NCCL, so forces agree to
    rounding, not bitwise: rigid body within 1e-10 after 500 steps, flexible bodies on the early trajectory (see above)."""
    if _ngpus() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    if not (_have(case, "LIFE_b200") and _have(case, "LIFE_ref")):
        pytest.skip("life_b200/host/_build/%s not built" % case)
    ref = _run(case, "LIFE_ref", str(tmp_path / "ref"))
    assert ref.returncode == 0, ref.stdout[-2000:]
    new = _run(case, "LIFE_b200", str(tmp_path / "b200"), LIFE_B200_GPUS=str(gpus), LIFE_B200_EXACT="1")
    assert new.returncode == 0, new.stdout[-2000:] + new.stderr[-2000:]
    assert "%d GPU(s)" % gpus in new.stderr
    differing, n_files = _diff_r(str(tmp_path / "ref"), str(tmp_path / "b200"))
    t_end, err = _compare(case, str(tmp_path / "ref"), str(tmp_path / "b200"))
    print("\n%s on %d GPUs, exact mode: %d files, differing %s; %s" % (case, gpus, n_files, differing, err))
    assert t_end == 500
    if case == "ChannelFlow":
        assert not differing, differing
    elif case == "Cylinder":
        bad = {k: v for k, v in err.items() if not v <= (1e-8 if k == "TotalForces.out" else K.TOL)}
        assert not bad, bad
    else:
        ta = R.read_table(str(tmp_path / "ref" / "Results" / "TotalForces.out"))
        tb = R.read_table(str(tmp_path / "b200" / "Results" / "TotalForces.out"))
        early = (ta[:, 0] > 0) & (ta[:, 0] <= 20)
        assert np.abs(tb[early, 1:] - ta[early, 1:]).max() <= 1e-7 * np.abs(ta[:, 1:]).max()


def test_program_refuses_to_run_without_a_gpu(tmp_path):
    """No CPU fallback: without a CUDA device the drop-in program stops through the reference's ERROR() (exit 99)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    case = "LidDrivenCavity"
    if not _have(case, "LIFE_b200"):
        pytest.skip("life_b200/host/_build not built")
    p = _run(case, "LIFE_b200", str(tmp_path / "run"))
    assert p.returncode == 99
    assert "no CPU path" in p.stdout or "life_create failed" in p.stdout


@pytest.mark.parametrize("case", K.EXAMPLES_LBM)
def test_oracle_matches_the_reference_program(case, tmp_path):
    """Pins the oracle at whole-program level: 500 steps of the unmodified reference executable (its Fluid.restart) against
    500 steps of oracle/life_oracle.c from the same initial state."""
    if not _have(case, "LIFE_ref"):
        pytest.skip("life_b200/host/_build/%s/LIFE_ref not built" % case)
    p = _run(case, "LIFE_ref", str(tmp_path / "ref"))
    assert p.returncode == 0
    a = R.read_fluid(str(tmp_path / "ref" / "Results" / "Restart" / "Fluid.restart"))
    g = K.golden(case)
    assert not int(g["wavy"])
    o = K.make_oracle(g)
    o.step(500)
    for name in ("rho", "u", "f"):
        err = K.rel_l2(o.get(name), a[name])
        assert err < 1e-12, (case, name, err)
