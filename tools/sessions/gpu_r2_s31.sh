#!/usr/bin/env bash
# Session 31: look-ahead kernel with the probe bit cleaned on the FMA pipe (one LOP3 less on the ALU pipe), A/B by env.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "look or glue or golden" > $OUT/r2_pytest_s31.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2_pytest_s31.log
for clean in 1 0; do for regs in 48 40; do
  PIRE_B200_LOOK_CLEAN=$clean PIRE_B200_LOOK_REGS=$regs timeout 300 python bench.py --workload glue10 --variant look --steps 20 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_look_clean${clean}_r${regs}.json 2> $OUT/r2_bench_look_clean${clean}_r${regs}.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/r2_bench_look_clean${clean}_r${regs}.json')); print('look clean=$clean regs=$regs', round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],4), d['parity']['mismatches'])
except Exception as e: print('failed', e); print(open('$OUT/r2_bench_look_clean${clean}_r${regs}.err').read()[-1500:])
PY
done; done
