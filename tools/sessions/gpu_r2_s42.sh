#!/usr/bin/env bash
# Session 42: the full GPU suite once more (new AUTO-shape test), ncu captures of the uniform prefix kernel, the counting
# kernel and the split kernel as they ship.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2_final_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2_final_pytest_gpu.log; tail -3 $OUT/r2_final_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:PrefixUniform -s 10 -c 1 -f -o $OUT/r2_prof_prefix_uniform_glue10 \
    python tools/gpu_prefix_exp.py 4194304 > $OUT/r2_ncu_prefix_uniform.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:CountKernel -s 2 -c 1 -f -o $OUT/r2_prof_count_hf_glue10 \
    python tools/gpu_count_exp.py 4194304 > $OUT/r2_ncu_count.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanSplit -s 3 -c 1 -f -o $OUT/r2_prof_full_utf8mixed_split2 \
    python bench.py --workload utf8mixed --variant pred --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --no-next --no-parity > $OUT/r2_ncu_full_split2.log 2>&1
ls -la $OUT/r2_prof_prefix_uniform_glue10.ncu-rep $OUT/r2_prof_count_hf_glue10.ncu-rep $OUT/r2_prof_full_utf8mixed_split2.ncu-rep
