#!/bin/bash
# threaded file transfers: parity again, fresh-file timings, whole programs
TAG=${1:-r01g}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest: device-fed files"; timeout 600 python -m pytest tests/test_gpu_output.py -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_output.txt
echo "== io bench"; timeout 900 python scripts/io_bench.py 4096 16384 2>&1 | grep -v "^Writing restart" | tee $OUT/${TAG}_io_bench.txt
echo "== whole programs"; bash scripts/gpu_programs.sh Cavity4096 InvertedFlag Honami; cp $OUT/programs_timing.txt $OUT/${TAG}_programs_timing.txt
