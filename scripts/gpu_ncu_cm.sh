#!/bin/bash
# ncu --set full capture of the central-moments bulk sweep at 16384^2 (2 launches), exported to csv / text on the box
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bulk -s 6 -c 2 -f -o $OUT/r01_bulk_cm_16384 \
    python bench.py --collision cm --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r01_ncu_cm.log 2>&1
tail -2 $OUT/r01_ncu_cm.log
ncu -i $OUT/r01_bulk_cm_16384.ncu-rep --page raw --csv > $OUT/r01_bulk_cm_16384_raw.csv 2>/dev/null
ncu -i $OUT/r01_bulk_cm_16384.ncu-rep --page details > $OUT/r01_bulk_cm_16384_details.txt 2>/dev/null
ls -la $OUT | grep cm_16384
