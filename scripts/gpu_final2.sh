#!/bin/bash
# last confirmation of the round: whole GPU suite + smoke
TAG=${1:-r01i}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
