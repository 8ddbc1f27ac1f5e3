#!/bin/bash
# 2 GPUs: slab parity after the refactor + the shared-file writes / reads
TAG=${1:-r01f}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -k "write_one_file or two_slabs" 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_multi.txt
