#!/bin/bash
TAG=${1:-r01t}
OUT=gpurun_out; mkdir -p $OUT
for n in 8 4; do
  echo "== strong scaling 32768^2 on $n GPUs"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --size 32768 --global-nx 32768 --steps 200 --warmup 10 > $OUT/${TAG}_strong_n$n.json 2> $OUT/${TAG}_strong_n$n.err
  grep "rank 0\]" $OUT/${TAG}_strong_n$n.err | tail -2; cat $OUT/${TAG}_strong_n$n.json
done
