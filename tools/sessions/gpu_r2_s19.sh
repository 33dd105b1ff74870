#!/usr/bin/env bash
# in-stream lines kernel with the segment size chosen on the device: parity, prose throughput with the reference beside it
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "lines" > $OUT/r2_pytest_lines.log 2>&1; echo "pytest lines exit $?"; tail -5 $OUT/r2_pytest_lines.log
timeout 900 python tools/gpu_prose_exp.py > $OUT/r2_prose_lines.log 2>&1; tail -3 $OUT/r2_prose_lines.log
PIRE_B200_LINES_KERNEL=1 PROSE_QUICK=1 timeout 600 python tools/gpu_prose_exp.py 2>&1 | tail -2
PROSE_QUICK=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:ScanTextKernel -s 2 -c 1 -o $OUT/r2_prof_text_kernel -f python tools/gpu_prose_exp.py > $OUT/r2_ncu_text.log 2>&1; tail -2 $OUT/r2_ncu_text.log
