#!/usr/bin/env bash
# in-stream lines kernel, data-driven version: parity, prose throughput, one ncu capture
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "lines" > $OUT/r2_pytest_lines.log 2>&1; echo "pytest lines exit $?"; tail -15 $OUT/r2_pytest_lines.log
PROSE_QUICK=1 timeout 600 python tools/gpu_prose_exp.py 2>&1 | tail -3
for seg in 1024 4096 8192; do
  echo "segment $seg"; PIRE_B200_TEXT_SEGMENT=$seg PROSE_QUICK=1 timeout 600 python tools/gpu_prose_exp.py 2>&1 | tail -2
done
PROSE_QUICK=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:ScanTextKernel -s 2 -c 1 -o $OUT/r2_prof_text_kernel -f python tools/gpu_prose_exp.py > $OUT/r2_ncu_text.log 2>&1; tail -3 $OUT/r2_ncu_text.log
