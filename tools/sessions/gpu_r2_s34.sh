#!/usr/bin/env bash
# Session 34: look-ahead step in 5.5 instructions (odd bytes probed dirty, even bytes clean): parity and timing.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "look or glue or golden or headline" > $OUT/r2_pytest_s34a.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2_pytest_s34a.log
PIRE_B200_LOOK_ILP=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "look or glue or golden or headline" > $OUT/r2_pytest_s34b.log 2>&1; echo "pytest ilp2 exit $?"; tail -2 $OUT/r2_pytest_s34b.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload glue10 --variant look --steps 20 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_$name.json 2> $OUT/r2_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/r2_bench_$name.json')); print('$name', round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],4), d['parity']['mismatches'])
except Exception as e: print('$name failed', e); print(open('$OUT/r2_bench_$name.err').read()[-1500:])
PY
}
run look55_b640 PIRE_B200_LOOK_BLOCK=640
run look55_b384 PIRE_B200_LOOK_BLOCK=384
run look55_b512r40 PIRE_B200_LOOK_REGS=40
run look55_ilp2_r80 PIRE_B200_LOOK_ILP=2 PIRE_B200_LOOK_ILP_REGS=80
run look55_ilp2_r72 PIRE_B200_LOOK_ILP=2 PIRE_B200_LOOK_ILP_REGS=72
