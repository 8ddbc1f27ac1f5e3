"""The run-time front end (life_b200/host/life_run.cpp): the case comes from a FILE instead of a compiled-in
params.h, the time loop runs through the C ABI on 1..N GPUs (one host thread each).  Checked against the unmodified reference
PROGRAM compiled for the same case (life_b200/host/_build/<case>/LIFE_ref): same protocol as tests/test_host_program.py —
500 steps, then the Results/ trees are compared: bit for bit in exact mode (BGK), within 1e-10 otherwise."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import cases as K
from tests import restartfile as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "life_b200", "host", "_build")
EXE = os.path.join(HOST, "life_run")
CASES = os.path.join(ROOT, "examples", "cases")


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _life_run(case, workdir, *overrides):
    os.makedirs(workdir, exist_ok=True)
    return subprocess.run([EXE, os.path.join(CASES, case + ".case")] + list(overrides), cwd=workdir, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=900)


def _reference(case, workdir):
    os.makedirs(workdir, exist_ok=True)
    p = subprocess.run([os.path.join(HOST, case, "LIFE_ref")], cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=900, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    assert p.returncode == 0, p.stdout[-2000:]


def test_life_run_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(EXE):
        pytest.skip("life_b200/host/_build/life_run not built")
    p = _life_run("ChannelFlow", str(tmp_path))
    assert p.returncode == 99 and "no CPU path" in p.stdout


def test_life_run_rejects_a_bad_case_file(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("life_b200/host/_build/life_run not built")
    bad = tmp_path / "bad.case"
    bad.write_text("Nx = 10\nNy = 10\nnu_p = 0.1\nomega = 1.0\ntStep = 0.1\n")
    p = subprocess.run([EXE, str(bad)], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert p.returncode == 99 and "exactly one of omega and tStep" in p.stdout
    bad.write_text("Nx = 10\nNy = 10\nnu_p = 0.1\nomega = 1.0\nWALL_TOP = eLid\n")
    p = subprocess.run([EXE, str(bad)], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert p.returncode == 99 and "not a lattice-site type" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [1, 2, 4])
@pytest.mark.parametrize("case,exact", [("ChannelFlow", 1), ("ChannelFlow", 0), ("LidDrivenCavity", 1), ("LidDrivenCavity", 0)])
def test_life_run_reproduces_the_reference_program(case, exact, gpus, tmp_path):
    if not (os.path.exists(EXE) and os.path.exists(os.path.join(HOST, case, "LIFE_ref"))):
        pytest.skip("life_b200/host/_build not built")
    if _ngpus() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    _reference(case, str(tmp_path / "ref"))
    p = _life_run(case, str(tmp_path / "run"), "exact=%d" % exact, "gpus=%d" % gpus)
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count("Time step ") == 51 and "FINISHED" in p.stdout
    a_dir, b_dir = tmp_path / "ref" / "Results", tmp_path / "run" / "Results"
    fluid = lambda d: sorted(f for f in os.listdir(d / "VTK") if f.startswith("Fluid."))
    assert fluid(a_dir) == fluid(b_dir) and len(fluid(a_dir)) == 11
    a, b = R.read_fluid(str(a_dir / "Restart" / "Fluid.restart")), R.read_fluid(str(b_dir / "Restart" / "Fluid.restart"))
    assert (a["t"], a["Nx"], a["Ny"]) == (b["t"], b["Nx"], b["Ny"]) == (500, a["Nx"], a["Ny"])
    if exact:
        # the reference's own protocol: diff -r (fluid files; the reference also writes Log.out, which holds timings)
        assert open(a_dir / "Restart" / "Fluid.restart", "rb").read() == open(b_dir / "Restart" / "Fluid.restart", "rb").read()
        for f in fluid(a_dir):
            assert open(a_dir / "VTK" / f, "rb").read() == open(b_dir / "VTK" / f, "rb").read(), f
    else:
        for name in ("rho", "u", "f"):
            err = K.rel_l2(b[name], a[name])
            assert err < K.TOL, (case, name, err)
        assert open(a_dir / "VTK" / "Fluid.0.vti", "rb").read() == open(b_dir / "VTK" / "Fluid.0.vti", "rb").read()


@pytest.mark.gpu
def test_life_run_restarts_from_its_own_restart_file(tmp_path):
    """500 + 500 steps with a restart in between == the reference run twice (testing/store-ref-data.sh:51-53), bit for bit."""
    case = "ChannelFlow"
    if not (os.path.exists(EXE) and os.path.exists(os.path.join(HOST, case, "LIFE_ref"))):
        pytest.skip("life_b200/host/_build not built")
    for _ in range(2):
        _reference(case, str(tmp_path / "ref"))
        p = _life_run(case, str(tmp_path / "run"), "exact=1")
        assert p.returncode == 0, p.stdout[-3000:]
    assert "continuing from time step 500" in p.stdout
    a = open(tmp_path / "ref" / "Results" / "Restart" / "Fluid.restart", "rb").read()
    b = open(tmp_path / "run" / "Results" / "Restart" / "Fluid.restart", "rb").read()
    assert R.read_fluid(str(tmp_path / "run" / "Results" / "Restart" / "Fluid.restart"))["t"] == 1000
    assert a == b
