#!/usr/bin/env bash
# Session 46: HalfFinalScanner counting with the look-ahead first pass (large automata, uniform batches): parity, then
# tools/gpu_count_exp.py with and without it on the same box.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "count or half_final" > $OUT/r2_pytest_s46.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r2_pytest_s46.log
for look in 1 0; do
  PIRE_B200_COUNT_LOOK=$look timeout 400 python tools/gpu_count_exp.py 4194304 > $OUT/r2_count_look$look.json 2> $OUT/r2_count_look$look.err
  python -c "
import json; d=json.load(open('$OUT/r2_count_look$look.json')); print('look=$look', {k: {m: round(v['GBps'],1) for m,v in d[k].items() if isinstance(v, dict) and 'GBps' in v} for k in ('hf_glue10','count_words5')})" || tail -5 $OUT/r2_count_look$look.err
done
