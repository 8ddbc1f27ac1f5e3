"""Build liblife_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python -m life_b200.build [--force] [--verbose]

Output: life_b200/lib/liblife_b200.so  (git-ignored; travels to the GPU box with the gpurun snapshot).
nvcc cross-compiles without a GPU.  Objects are rebuilt only when a source or header is newer.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "liblife_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["api.cu", "lbm_bulk.cu", "lbm_boundary.cu", "lbm_io.cu", "lbm_file.cu", "halo.cu", "ibm.cu", "ibm_eps.cu", "fem.cu", "nccl_dyn.cu", "membw.cu", "lbm_small.cu"]
# second compilation of the step kernels in the reference's operation order (cfg.exact): no FMA contraction, namespace life::exact
EXACT_FLAGS = ["-DLIFE_EXACT", "-fmad=false"]
VARIANTS = [(s, s.replace(".cu", ".o"), []) for s in SOURCES] + \
           [(s, s.replace(".cu", "_exact.o"), EXACT_FLAGS) for s in ("lbm_bulk.cu", "lbm_boundary.cu", "lbm_small.cu")]

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
# sm_100a only (B200); -lineinfo so ncu's source page maps to these files.  The host compiler is the system g++.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-O3,-Wall", "-I", INCLUDE]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers.append(os.path.join(INCLUDE, "life_b200.h"))
    jobs = []
    me = os.path.abspath(__file__)
    for s, o, vflags in VARIANTS:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, o)
        if force or _newer(obj, [src, me] + headers):
            jobs.append([NVCC] + NVCC_FLAGS + vflags + list(extra_flags) + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr, flush=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, o) for _, o, _ in VARIANTS]
    if force or jobs or _newer(LIB, objs):
        run([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
