#!/bin/bash
# Whole-program timing of the example cases: unmodified reference program vs the drop-in program, 500-step protocol.
OUT=gpurun_out; mkdir -p $OUT; B=$PWD/life_b200/host/_build
export OPENBLAS_NUM_THREADS=1
for c in "$@"; do
  for mode in ref b200:0 b200:3; do
    exe=LIFE_${mode%%:*}; eps=${mode##*:}
    d=$(mktemp -d); [ -d $B/$c/input ] && cp -r $B/$c/input $d/
    ( cd $d; LIFE_B200_DEVICE_EPSILON=$eps $B/$c/$exe > log.txt 2> err.txt; echo "== $c $mode: $(grep -o 'Simulation took [0-9.]* seconds' log.txt)"; grep "life_b200" err.txt )
    rm -rf $d
  done
done 2>&1 | tee $OUT/examples_timing.txt
