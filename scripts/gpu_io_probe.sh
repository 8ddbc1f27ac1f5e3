#!/bin/bash
# where do the output calls of the small FSI programs spend their time?  (report line of life_host.cpp, 3 runs each)
OUT=gpurun_out; mkdir -p $OUT; B=$PWD/life_b200/host/_build
export OPENBLAS_NUM_THREADS=1
for c in TurekHron PELskin ChannelFlow; do
  for io in 0 1 0 1 0 1; do
    d=$(mktemp -d); [ -d $B/$c/input ] && cp -r $B/$c/input $d/
    ( cd $d; LIFE_B200_HOST_IO=$io $B/$c/LIFE_b200 > log.txt 2> err.txt; echo "== $c host_io=$io: $(grep -o 'Simulation took [0-9.]* seconds' log.txt)"; grep "wall" err.txt | sed -e 's/.*inside life_step/inside life_step/' )
    rm -rf $d
  done
done 2>&1 | tee $OUT/io_probe.txt
