#!/usr/bin/env bash
# in-stream lines kernel: parity, prose throughput against the pulling-lanes kernel, memcheck
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "lines" > $OUT/r2_pytest_lines.log 2>&1; echo "pytest lines exit $?"; tail -15 $OUT/r2_pytest_lines.log
for k in 1 2; do
  echo "PIRE_B200_LINES_KERNEL=$k"
  PIRE_B200_LINES_KERNEL=$k PROSE_QUICK=1 timeout 600 python tools/gpu_prose_exp.py 2>&1 | tail -3
done
for seg in 512 2048 4096; do
  echo "segment $seg"; PIRE_B200_TEXT_SEGMENT=$seg PROSE_QUICK=1 timeout 600 python tools/gpu_prose_exp.py 2>&1 | tail -2
done
timeout 900 python tools/gpu_prose_exp.py > $OUT/r2_prose_lines.log 2>&1; tail -3 $OUT/r2_prose_lines.log
