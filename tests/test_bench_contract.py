"""bench.py's line format (the driver parses it): checked on CPU through the reference arm, which needs no GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    from oracle import refharness
    if not refharness.available("syn_bgk"):
        pytest.skip("oracle/_ref not built")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["metric"].startswith("MLUPS") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_every_core_under_torchrun():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the reference leg must not inherit that (round 1's
    multi-GPU reference numbers were single threaded).  The team size is the one the reference's own counter observed."""
    from oracle import refharness
    if not refharness.available("syn_bgk"):
        pytest.skip("oracle/_ref not built")
    cores = len(os.sched_getaffinity(0))
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == cores
    assert all(s_["threads"] == cores for s_ in d["cpu_baseline"]["samples"])
    assert {s_["lattice"] for s_ in d["cpu_baseline"]["samples"]} == {"8192x8192", "4096x4096"}
    # the line says what it timed: the CPU sample's own lattice, not the GPU arm's
    assert d["config"]["Nx"] == 8192 and "CPU SAMPLE" in d["config"]["workload"] and d["config"]["gpu_arm_lattice"] == "32768x16384"


def test_fixture_state_rebuilds_the_golden_initial_state_without_the_oracle():
    """bench.py's untimed multi-slab parity check configures its cases from the committed fixtures alone (tests/fixture_state.py);
    that must be the same configuration and state the oracle-based tests start from."""
    import numpy as np
    from tests import cases as K, fixture_state as FS
    for case in ("t_periodic_cm", "t_periodic_bgk", "t_womersley", "t_convective", "Cylinder", "Honami"):
        g = K.golden(case)
        o = K.make_oracle(g)
        a = K.life_config(o.params, o)
        kw = FS.config_kwargs(FS.load(case))
        for k, v in kw.items():
            assert getattr(a, k) == v, (case, k, getattr(a, k), v)
        f, rho, u, fxy, u_in, rho_in = FS.initial_state(FS.load(case))
        for name, arr in (("f", f), ("rho", rho), ("u", u), ("force_xy", fxy), ("u_in", u_in), ("rho_in", rho_in)):
            assert np.array_equal(arr, o.get(name)), (case, name)


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_on_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2"], cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode != 0
    assert "no CPU path" in p.stderr + p.stdout
