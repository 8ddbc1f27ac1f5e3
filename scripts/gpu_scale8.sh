#!/bin/bash
TAG=${1:-r01s}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt; free -g >> $OUT/${TAG}_gpus.txt
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 200 --warmup 10 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
grep -v "^W1017\|^\*\*\*\|OMP_NUM" $OUT/${TAG}_bench_n$N.err | tail -12; cat $OUT/${TAG}_bench_n$N.json
echo "== parity on $N slabs"
bash scripts/gpu_mp_debug.sh $N t_periodic_cm Honami PELskin InvertedFlag ChannelFlow
if [ $N -ge 8 ]; then
echo "== bench N=4"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 4 --steps 200 --warmup 10 > $OUT/${TAG}_bench_n4.json 2> $OUT/${TAG}_bench_n4.err
cat $OUT/${TAG}_bench_n4.json
fi
