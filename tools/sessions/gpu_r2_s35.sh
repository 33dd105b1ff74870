#!/usr/bin/env bash
# Session 35: uniform prefix kernel (no ring, exit filter): parity and timing; default look-ahead shape sanity.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefix or suffix or uniform_bodies" > $OUT/r2_pytest_s35.log 2>&1; echo "pytest prefix exit $?"; tail -2 $OUT/r2_pytest_s35.log
PIRE_B200_PREFIX_PRED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefix or suffix or uniform_bodies" > $OUT/r2_pytest_s35b.log 2>&1; echo "pytest prefix pred=1 exit $?"; tail -2 $OUT/r2_pytest_s35b.log
for cfg in "1 640" "0 640" "1 384" "1 512" "0 384"; do
  set -- $cfg
  PIRE_B200_PREFIX_PRED=$1 PIRE_B200_PREFIX_BLOCK=$2 timeout 300 python tools/gpu_prefix_exp.py 4194304 > $OUT/r2_prefix_pred$1_b$2.json 2> $OUT/r2_prefix_pred$1_b$2.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/r2_prefix_pred$1_b$2.json'))
    print('prefix pred=$1 block=$2', {k:{m:round(v[m]['GBps'],1) for m in v} for k,v in d.items() if isinstance(v,dict)})
except Exception as e: print('failed', e); print(open('$OUT/r2_prefix_pred$1_b$2.err').read()[-1200:])
PY
done
PIRE_B200_NO_UNIFORM_BODY=1 timeout 300 python tools/gpu_prefix_exp.py 4194304 > $OUT/r2_prefix_ring.json 2>&1; python -c "
import json; d=json.load(open('$OUT/r2_prefix_ring.json')); print('prefix ring', {k:{m:round(v[m]['GBps'],1) for m in v} for k,v in d.items() if isinstance(v,dict)})"
timeout 300 python bench.py --workload glue10 --steps 20 --warmup 3 --no-e2e --no-cpu --no-configs > $OUT/r2_bench_glue10_auto_s35.json 2> $OUT/r2_bench_glue10_auto_s35.err
python -c "
import json; d=json.load(open('$OUT/r2_bench_glue10_auto_s35.json')); print('glue10 auto', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['config']['variant_ms'], d['parity']['mismatches'], d.get('next_rows'))" || tail -5 $OUT/r2_bench_glue10_auto_s35.err
