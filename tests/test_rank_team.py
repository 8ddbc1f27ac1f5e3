"""Host logic of the multi-GPU host programs that needs no GPU: the one-thread-per-rank team (life_b200/host/rank_team.h) under
ThreadSanitizer, and the build rule of the exact-mode objects (no FMA contraction in their PTX)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_team_under_thread_sanitizer(tmp_path):
    exe = str(tmp_path / "rank_team_test")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-I", os.path.join(ROOT, "life_b200", "host"),
                           os.path.join(ROOT, "tests", "native", "rank_team_test.cpp"), "-o", exe, "-lpthread"])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == "OK", p.stdout[-2000:] + p.stderr[-3000:]
    assert "ThreadSanitizer" not in p.stderr, p.stderr[-3000:]


@pytest.mark.parametrize("src", ["lbm_bulk.cu", "lbm_boundary.cu", "lbm_small.cu"])
def test_exact_objects_contain_no_fused_multiply_add(src, tmp_path):
    """cfg.exact relies on the second compilation of the step kernels (life_b200/build.py: EXACT_FLAGS) rounding every product and
    sum on its own, like the reference's g++ build for baseline x86-64.  The PTX of that compilation must not contain a single
    fma; division and square root stay the IEEE `div.rn.f64` / `sqrt.rn.f64`."""
    from life_b200 import build as B
    nvcc = B.NVCC if os.path.exists(B.NVCC) else shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not available")
    assert B.EXACT_FLAGS == ["-DLIFE_EXACT", "-fmad=false"]
    ptx = str(tmp_path / (src + ".ptx"))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "-I", B.INCLUDE] + B.EXACT_FLAGS +
                          ["-ptx", os.path.join(B.CSRC, src), "-o", ptx], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    text = open(ptx).read()
    assert ".entry" in text and "life5exact" in text          # kernels of namespace life::exact
    assert "fma." not in text, [ln for ln in text.splitlines() if "fma." in ln][:5]
    assert "div.rn.f64" in text
    assert "div.approx" not in text and "div.full" not in text and "rcp.approx" not in text
    # and the default compilation of the same source does contract (otherwise this test would prove nothing)
    ptx2 = str(tmp_path / (src + ".fast.ptx"))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "-I", B.INCLUDE, "-ptx",
                           os.path.join(B.CSRC, src), "-o", ptx2], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert "fma.rn.f64" in open(ptx2).read()
