#!/bin/bash
# Whole-program timing of the flexible-body examples (500-step protocol): the unmodified reference program vs the drop-in with the
# reference's host FEM (default), with the structural solver on the device (LIFE_B200_DEVICE_FEM=1) and with the whole sub-iteration
# loop resident on the device (=2).  Prints wall time, time inside objectKernel and microseconds per sub-iteration.
# Usage: scripts/fsi_timing.sh CASE...   (cases under life_b200/host/_build)
OUT=gpurun_out; mkdir -p $OUT; B=$PWD/life_b200/host/_build
export OPENBLAS_NUM_THREADS=1
for c in "$@"; do
  for mode in ref host-fem device-fem resident; do
    exe=LIFE_b200; lvl=0
    case $mode in ref) exe=LIFE_ref;; device-fem) lvl=1;; resident) lvl=2;; esac
    d=$(mktemp -d); [ -d $B/$c/input ] && cp -r $B/$c/input $d/
    ( cd $d; LIFE_B200_DEVICE_FEM=$lvl timeout 900 $B/$c/$exe > log.txt 2> err.txt
      echo "== $c $mode: $(grep -o 'Simulation took [0-9.]* seconds' log.txt)"
      grep -E "objectKernel|steady state|wall " err.txt | sed 's/^/     /' )
    rm -rf $d
  done
done 2>&1 | tee $OUT/fsi_timing.txt
