"""examples/cavity.c — the C ABI used from a plain C program — builds against include/life_b200.h and liblife_b200.so, refuses to
run without a B200 (no CPU fallback), and on a B200 produces the reference's file formats."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cavity_exe(tmp_path_factory, lib_built):
    exe = str(tmp_path_factory.mktemp("example") / "cavity")
    lib = os.path.join(ROOT, "life_b200", "lib")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cavity.c"),
                           "-L", lib, "-llife_b200", "-Wl,-rpath," + lib, "-lm", "-o", exe])
    return exe


def test_example_builds_and_refuses_to_run_without_a_gpu(cavity_exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([cavity_exe, "64", "10", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 99 and "no CPU path" in p.stderr


@pytest.mark.gpu
def test_example_runs_and_writes_the_reference_formats(cavity_exe, tmp_path):
    from oracle import fluidfiles as F
    N, steps = 257, 40
    p = subprocess.run([cavity_exe, str(N), str(steps), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("max |u|") == 10
    rst = F.read_restart(str(tmp_path / "Fluid.restart"))
    assert (rst["t"], rst["Nx"], rst["Ny"]) == (steps, N, N)
    vti = F.read_vti(str(tmp_path / ("Fluid.%d.vti" % steps)), N, N)
    assert abs(float(vti["density"].mean()) - 1.0) < 1e-3            # Drho = 1: the file holds rho
    assert 0.0 < float(abs(vti["velocity"][:, :, 0]).max()) <= 1.0 + 1e-9   # lid speed 0.1 lattice units = 1 m/s
    assert os.path.exists(tmp_path / ("Fluid.%d.vti" % (steps // 2)))
