#!/bin/bash
# device-fed files: new parity tests first (fast feedback), then the whole GPU suite, then the I/O timings
TAG=${1:-r01d}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
echo "== pytest: device-fed files"; timeout 600 python -m pytest tests/test_gpu_output.py -x -q -s 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_output.txt
echo "== pytest: whole GPU suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== io bench"; timeout 600 python scripts/io_bench.py 4096 8192 2>&1 | tee $OUT/${TAG}_io_bench.txt
