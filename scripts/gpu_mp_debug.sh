#!/bin/bash
# Usage: bash scripts/gpu_mp_debug.sh NPROC "case [steps]" ...   — runs tests/mp_parity.py under torchrun, keeps the rank lines
OUT=gpurun_out; mkdir -p $OUT
N=$1; shift
for c in "$@"; do
  echo "== $c"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mp_parity.py $c 2>&1 | grep "^\[rank\|Error\|error:\|assert" | head -20
done 2>&1 | tee $OUT/mp_debug.txt
