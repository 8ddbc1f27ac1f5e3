#!/bin/bash
# Whole-program timing: the unmodified reference program vs the drop-in program with device-fed files (default) and with the
# host-mirror I/O (LIFE_B200_HOST_IO=1).  Usage: gpu_programs.sh CASE...   (cases under life_b200/host/_build)
OUT=gpurun_out; mkdir -p $OUT; B=$PWD/life_b200/host/_build
export OPENBLAS_NUM_THREADS=1
for c in "$@"; do
  for mode in ref b200:device-files b200:host-io; do
    exe=LIFE_${mode%%:*}; io=0; [ "${mode##*:}" = host-io ] && io=1
    d=$(mktemp -d); [ -d $B/$c/input ] && cp -r $B/$c/input $d/
    ( cd $d; LIFE_B200_HOST_IO=$io timeout 900 $B/$c/$exe > log.txt 2> err.txt
      echo "== $c $mode: $(grep -o 'Simulation took [0-9.]* seconds' log.txt)   MLUPS line: $(grep 'MLUPS' log.txt | tail -1)"
      grep "life_b200" err.txt; ls -la Results/VTK 2>/dev/null | tail -2; ls -la Results/Restart 2>/dev/null | tail -2 )
    rm -rf $d
  done
done 2>&1 | tee $OUT/programs_timing.txt
