#!/bin/bash
# One GPU-box session: parity tests, smoke, contract bench, kernel-variant micro-bench, ncu launch list + full capture.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/${TAG}_nvidia_smi.txt 2>&1
nproc > $OUT/${TAG}_host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/${TAG}_host.txt; free -g >> $OUT/${TAG}_host.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
echo "== bench cm"; timeout 900 python bench.py --steps 100 --warmup 10 --collision cm --no-cpu-baseline > $OUT/${TAG}_bench_cm.json 2> $OUT/${TAG}_bench_cm.err; cat $OUT/${TAG}_bench_cm.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tee $OUT/${TAG}_bench_reference.json
echo "== quick"; timeout 600 python scripts/quick_bench.py 16384 20 2>&1 | tee $OUT/${TAG}_quick_16384.txt
timeout 300 python scripts/quick_bench.py 4096 50 2>&1 | tee $OUT/${TAG}_quick_4096.txt
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bulk -s 6 -c 2 -f -o $OUT/${TAG}_bulk \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
