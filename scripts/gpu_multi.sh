#!/bin/bash
# Multi-GPU session: full GPU test suite (incl. 2-slab parity and the host drop-in program) + the contract bench at N GPUs.
# Usage (under gpurun --gpus N):  bash scripts/gpu_multi.sh TAG N [steps]
TAG=${1:-r01m}; N=${2:-2}; STEPS=${3:-100}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt; free -g >> $OUT/${TAG}_gpus.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
for n in 1 $N; do
  echo "== bench N=$n"
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps $STEPS --warmup 10 > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
  fi
  tail -4 $OUT/${TAG}_bench_n$n.err; cat $OUT/${TAG}_bench_n$n.json
done
