#!/bin/bash
TAG=${1:-r01b}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest lbm"; timeout 900 python -m pytest tests/test_gpu_lbm.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest_lbm.txt
echo "== tune"; timeout 600 python scripts/tune_bulk.py 16384 20 2>&1 | tee $OUT/${TAG}_tune.txt
echo "== bench (whole slab)"; timeout 900 python bench.py --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -2 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
echo "== bench (streamed host image)"; timeout 900 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --host-chunk-columns 1024 > $OUT/${TAG}_bench_stream.json 2> $OUT/${TAG}_bench_stream.err; tail -2 $OUT/${TAG}_bench_stream.err; cat $OUT/${TAG}_bench_stream.json
