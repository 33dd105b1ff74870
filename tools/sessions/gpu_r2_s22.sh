#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ScanSplitKernel -s 3 -c 1 -o $OUT/r2_prof_split_kernel -f python bench.py --workload utf8mixed --variant pred --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_ncu_split.log 2>&1; tail -2 $OUT/r2_ncu_split.log
