#!/bin/bash
# ncu evidence of a round (B200_PROFILING.md recipe), one GPU: launch list of the bench command, full captures of the sweep kernels,
# launch list of the marker kernels.  Usage: scripts/profile_round.sh r02
R=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-parity-check"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${R}_launches_bench16384.csv $B > $OUT/${R}_launches_bench.log 2>&1
cap() {   # name, kernel regex, extra bench flags
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 6 -c 2 -f -o $OUT/${R}_$1 $B $3 > $OUT/${R}_$1.log 2>&1
  ncu -i $OUT/${R}_$1.ncu-rep --page details > $OUT/${R}_$1_ncu_details.txt 2>/dev/null
  ncu -i $OUT/${R}_$1.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
keep=[i for i,h in enumerate(rows[0]) if any(k in h for k in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__throughput.avg.pct','smsp__cycles_active.avg','launch__grid_size','launch__block_size'))]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in keep if i < len(r)])
" > $OUT/${R}_$1_ncu_raw_selected.csv
  rm -f $OUT/${R}_$1.ncu-rep
}
cap bulk_bgk_16384 k_bulk_shuffle ""
cap bulk_cm_16384 k_bulk_shuffle "--collision cm"
cap bulk_quad_16384 k_bulk_quad "--kernel 4"
cap bulk_shift_16384 k_bulk_shift "--inplace"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${R}_launches_ibm_trace_replay.csv \
    python -m pytest tests/test_gpu_ibm.py -q -k "trace_replay and ordered and (PELskin or Honami)" > $OUT/${R}_launches_ibm.log 2>&1
ls -la $OUT | tail -20
