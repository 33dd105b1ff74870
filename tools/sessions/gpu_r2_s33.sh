#!/usr/bin/env bash
# Session 33: look-ahead kernel with two strings per lane (ILP 2): parity, register budgets / CTA shapes.
set -u
OUT=gpurun_out
mkdir -p $OUT
PIRE_B200_LOOK_ILP=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "look or glue or golden or headline" > $OUT/r2_pytest_s33.log 2>&1; echo "pytest ilp2 exit $?"; tail -2 $OUT/r2_pytest_s33.log
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
run look2_r72 PIRE_B200_LOOK_ILP=2 PIRE_B200_LOOK_ILP_REGS=72
run look2_r80 PIRE_B200_LOOK_ILP=2 PIRE_B200_LOOK_ILP_REGS=80
run look2_r64 PIRE_B200_LOOK_ILP=2 PIRE_B200_LOOK_ILP_REGS=64
run look2_r72_b416 PIRE_B200_LOOK_ILP=2 PIRE_B200_LOOK_ILP_REGS=72 PIRE_B200_LOOK_BLOCK=416
run look_b640 PIRE_B200_LOOK_BLOCK=640
