"""First device run of the structural solver: executes the script of tests/test_gpu_fem.py in-process for one case.
    python scripts/gpu_fem_probe.py CASE STEPS"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
from tests.test_gpu_fem import SCRIPT  # noqa: E402

exec(SCRIPT % dict(root=ROOT, case=sys.argv[1], steps=int(sys.argv[2])))
