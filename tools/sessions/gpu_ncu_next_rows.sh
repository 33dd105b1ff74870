#!/usr/bin/env bash
# ncu --set full captures of the kernels behind the "next" rows and the ragged batch: generic (length-binned
# utf8mixed), counting (hf_glue10) and prefix (glue10 longest).  One launch each, small inputs.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanGeneric -s 4 -c 1 -f -o $OUT/prof_utf8mixed_binned \
    python bench.py --workload utf8mixed --strings 320000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant plain > $OUT/ncu_full_mixed.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:CountKernel -s 2 -c 1 -f -o $OUT/prof_count_hf_glue10 \
    python tools/gpu_count_exp.py 1048576 > $OUT/ncu_full_count.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:PrefixKernel -s 6 -c 1 -f -o $OUT/prof_prefix_glue10 \
    python tools/gpu_prefix_exp.py 1048576 > $OUT/ncu_full_prefix.log 2>&1
ls -la $OUT/*.ncu-rep
