#!/usr/bin/env bash
# Session 36: uniform prefix kernel: all-final fast path, IDP byte extraction, CTA shapes.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefix or suffix or uniform_bodies" > $OUT/r2_pytest_s36.log 2>&1; echo "pytest prefix exit $?"; tail -2 $OUT/r2_pytest_s36.log
PIRE_B200_PREFIX_IDP=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefix or suffix or uniform_bodies" > $OUT/r2_pytest_s36b.log 2>&1; echo "pytest prefix idp exit $?"; tail -2 $OUT/r2_pytest_s36b.log
for cfg in "0 640" "1 640" "0 512" "1 384"; do
  set -- $cfg
  PIRE_B200_PREFIX_IDP=$1 PIRE_B200_PREFIX_BLOCK=$2 timeout 300 python tools/gpu_prefix_exp.py 4194304 > $OUT/r2_prefix_idp$1_b$2.json 2> $OUT/r2_prefix_idp$1_b$2.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/r2_prefix_idp$1_b$2.json'))
    print('prefix idp=$1 block=$2', {k:{m:round(v[m]['GBps'],1) for m in v} for k,v in d.items() if isinstance(v,dict)})
except Exception as e: print('failed', e); print(open('$OUT/r2_prefix_idp$1_b$2.err').read()[-1200:])
PY
done
