#!/bin/bash
TAG=${1:-r01c}
OUT=gpurun_out; mkdir -p $OUT
echo "== ibm bench"; timeout 900 python scripts/ibm_bench.py 4096 20 2>&1 | tee $OUT/${TAG}_ibm_bench.txt
echo "== bench default"; timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -2 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
echo "== bench 32768^2 on one GPU"; timeout 900 python bench.py --size 32768 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_32768.json 2> $OUT/${TAG}_bench_32768.err; tail -3 $OUT/${TAG}_bench_32768.err; cat $OUT/${TAG}_bench_32768.json
echo "== ncu launch list: bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "== ncu launch list: PELskin + Honami trace replay (ordered spread)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_ibm.csv \
    python -m pytest tests/test_gpu_ibm.py -q -k "trace_replay and ordered and (PELskin or Honami)" > $OUT/${TAG}_ncu_ibm.log 2>&1
tail -2 $OUT/${TAG}_ncu_ibm.log
