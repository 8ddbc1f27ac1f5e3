#!/bin/bash
# end-of-session confirmation: whole GPU suite, smoke, default bench (both arms), whole-program timings
TAG=${1:-r01h}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; cat $OUT/${TAG}_bench_reference.json
echo "== bench (default)"; timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -2 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
echo "== whole programs"; bash scripts/gpu_programs.sh Cavity4096 LidDrivenCavity ChannelFlow Cylinder TurekHron InvertedFlag Honami PELskin > /dev/null 2>&1; cp $OUT/programs_timing.txt $OUT/${TAG}_programs_timing.txt; grep "^==" $OUT/${TAG}_programs_timing.txt
