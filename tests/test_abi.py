"""The drop-in boundary without a GPU: liblife_b200.so builds for sm_100a, loads, and exports exactly the entry points
include/life_b200.h declares; argument checking works; and creation fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "life_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(life_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from life_b200 import capi
    assert _declared() == sorted(capi.EXPORTS)


def test_library_exports_every_symbol(lib_built):
    from life_b200 import capi
    lib = C.CDLL(lib_built)
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.life_abi_version() == capi.ABI_VERSION == 2


def test_struct_layout_matches_header(lib_built):
    """sizeof(life_config) as the C compiler sees it == the ctypes mirror."""
    import subprocess
    import tempfile
    from life_b200 import capi
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "life_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", '
                             'sizeof(life_config), offsetof(life_config, omega), offsetof(life_config, stream), '
                             'offsetof(life_config, kernel));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) == C.sizeof(capi.Config)
    assert int(out[1]) == capi.Config.omega.offset
    assert int(out[2]) == capi.Config.stream.offset
    assert int(out[3]) == capi.Config.kernel.offset


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device life_create must fail with LIFE_E_CUDA and say why (never run on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from life_b200 import capi
    cfg = capi.Config(Nx=16, Ny=16, omega=1.0, Dx=1.0, Dt=1.0, Dm=1.0)
    with pytest.raises(capi.LifeError) as e:
        capi.Context(cfg)
    assert e.value.code == capi.E_CUDA
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_argument_checks(lib_built):
    from life_b200 import capi
    L = capi.load()
    h = C.c_void_p()
    bad = capi.Config(Nx=16, Ny=16, omega=1.0, Dx=1.0, Dt=1.0, Dm=1.0)
    bad.abi_version = 99
    assert L.life_create(C.byref(bad), C.byref(h)) == capi.E_ARG
    assert b"ABI" in L.life_last_error(None)
    bad = capi.Config(Nx=2, Ny=16, omega=1.0, Dx=1.0, Dt=1.0, Dm=1.0)
    assert L.life_create(C.byref(bad), C.byref(h)) == capi.E_ARG
    bad = capi.Config(Nx=16, Ny=16, omega=1.0, Dx=1.0, Dt=1.0, Dm=1.0, wall_left=7)
    assert L.life_create(C.byref(bad), C.byref(h)) == capi.E_ARG
    assert L.life_step(None, 1) == capi.E_ARG


def test_product_does_not_touch_the_oracle():
    """Nothing under life_b200/ or include/ may import, link, include or execute oracle/ (checker only) — sources, headers AND the
    build files (the host Makefile used to borrow two shim headers from oracle/: VERDICT round 1, hygiene 13)."""
    for base in ("life_b200", "include", "examples"):
        for dp, dirs, files in os.walk(os.path.join(ROOT, base)):
            dirs[:] = [d for d in dirs if d not in ("_build", "lib", "__pycache__")]
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".c", ".in", ".case")) or fn == "Makefile":
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    assert "oracle" not in txt.lower() or fn == "capi.py" and False, os.path.join(dp, fn)


def test_file_entry_points_reject_bad_arguments_without_a_device(lib_built):
    """The device-fed file entry points validate their arguments before touching CUDA (no context can exist on this machine)."""
    import ctypes as C
    from life_b200 import capi
    L = capi.load()
    assert L.life_write_vtk(None, b"/tmp/x.vti", 1.0, 0.0, capi.IO_SYNC) == capi.E_ARG
    assert L.life_write_restart(None, b"/tmp/x.restart", 1, capi.IO_ASYNC) == capi.E_ARG
    assert L.life_read_restart(None, b"/tmp/x.restart", None, None, None, None) == capi.E_ARG
    assert L.life_io_wait(None) == capi.E_ARG
    assert L.life_io_set_staging(None, 1 << 20) == capi.E_ARG
    busy = C.c_int32(7)
    assert L.life_io_busy(None, C.byref(busy)) == capi.E_ARG
    assert L.life_io_stats(None, None, None, None) == capi.E_ARG
    assert capi.E_IO == 7 and capi.IO_SYNC == 0 and capi.IO_ASYNC == 1


def test_fem_body_struct_layout_matches_the_header(tmp_path):
    """ctypes mirror of struct life_fem_body == the C layout (offsets printed by a program compiled against include/life_b200.h)."""
    import subprocess
    from life_b200 import capi
    fields = [n for n, _ in capi.FemBody._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "life_b200.h"\nint main(void) {\n' +
                   "".join('  printf("%%zu\\n", offsetof(life_fem_body, %s));\n' % n for n in fields) +
                   '  printf("%zu\\n", sizeof(life_fem_body));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert out[:-1] == [getattr(capi.FemBody, n).offset for n in fields]
    assert out[-1] == __import__("ctypes").sizeof(capi.FemBody)
