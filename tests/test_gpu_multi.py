"""Slab-decomposed runs on 2 (or more) B200s of one box: every rank's slab must match the oracle / the compiled reference's
fixtures exactly as the single-GPU run does.  Skipped on a 1-GPU box (the gloo tests in test_dist_gloo.py cover the host logic)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(n, case, steps=None, outdir=None, inplace=False):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "tests", "mp_parity.py"), case]
    if steps or outdir:
        cmd.append(str(steps or 0))
    if outdir:
        cmd.append(str(outdir))
    env = dict(os.environ, LIFE_TEST_INPLACE="1" if inplace else "0")
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)


# walls + corners, pressure/velocity ends, periodic x ring, y-periodic wrap at the faces, free slip, convective outlet
# (last three columns on the last rank), Womersley forcing, central moments
LBM_CASES = ["LidDrivenCavity", "ChannelFlow", "t_periodic_bgk", "t_periodic_cm", "t_yperiodic", "t_freeslip_cm",
             "t_convective", "t_womersley", "t_velocity_box", "t_pressure_left"]
# bodies: markers whose 3x3 supports straddle a slab face (Honami: 128 filaments along x; PELskin: periodic ring)
IBM_CASES = ["Cylinder", "TurekHron", "InvertedFlag", "Honami", "PELskin"]


@pytest.mark.parametrize("case", LBM_CASES + IBM_CASES)
def test_two_slabs_match(case):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    p = _launch(2, case)
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count(": ok") == 2, p.stdout[-3000:]


@pytest.mark.parametrize("n", [3, 4, 8])
@pytest.mark.parametrize("case", ["t_periodic_cm", "Honami"])
def test_more_slabs_match(case, n):
    if _ngpus() < n:
        pytest.skip("needs %d GPUs" % n)
    p = _launch(n, case)
    assert p.returncode == 0, p.stdout[-3000:]


@pytest.mark.parametrize("n", [2, 4])
@pytest.mark.parametrize("case", ["ChannelFlow", "t_periodic_cm", "Cylinder"])
def test_slabs_write_one_file_together(case, n, tmp_path):
    """life_write_vtk / life_write_restart with nranks > 1: every rank writes its byte ranges of the one shared file (sync and
    async), bytes identical to the reference format of the gathered state; life_read_restart returns each rank its slab."""
    if _ngpus() < n:
        pytest.skip("needs %d GPUs" % n)
    p = _launch(n, case, outdir=tmp_path)
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count(": ok") == n, p.stdout[-3000:]


@pytest.mark.parametrize("case", ["ChannelFlow", "t_periodic_cm", "t_yperiodic", "t_convective", "t_womersley", "Honami", "PELskin"])
def test_two_slabs_match_with_the_inplace_layout(case, tmp_path):
    """cfg.inplace across slabs: the halo travels through contiguous staging buffers (the planes are circular), the shared files are
    written from the shifted layout."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    p = _launch(2, case, outdir=str(tmp_path) if case in ("ChannelFlow", "t_periodic_cm") else None, inplace=True)
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count(": ok") == 2, p.stdout[-3000:]
