#!/bin/bash
# after the force_ibm / set_markers fix: the touched suites, I/O timings at the headline size, launch list of the file kernels
TAG=${1:-r01e}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest: files, markers, whole programs"
timeout 900 python -m pytest tests/test_gpu_output.py tests/test_gpu_ibm.py tests/test_host_program.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 | tee $OUT/${TAG}_pytest.txt
echo "== io bench 16384"; timeout 600 python scripts/io_bench.py 16384 2>&1 | tee $OUT/${TAG}_io_bench_16384.txt
echo "== ncu launch list: io bench 4096"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_io.csv \
    python scripts/io_bench.py 4096 > $OUT/${TAG}_ncu_io.log 2>&1
tail -3 $OUT/${TAG}_ncu_io.log
