#!/bin/bash
# default settings (device-fed files, epsilon assembly on the device + host LAPACK): the three UNI_EPSILON / many-body programs
OUT=gpurun_out; mkdir -p $OUT; B=$PWD/life_b200/host/_build
export OPENBLAS_NUM_THREADS=1
for c in PELskin TurekHron Honami; do
  d=$(mktemp -d); cp -r $B/$c/input $d/
  ( cd $d; $B/$c/LIFE_b200 > log.txt 2> err.txt; echo "== $c default: $(grep -o 'Simulation took [0-9.]* seconds' log.txt)"; grep "life_b200" err.txt )
  rm -rf $d
done 2>&1 | tee $OUT/programs_default_eps.txt
