#!/usr/bin/env bash
# Session 30: input-path microbenchmark (TMA / TMEM next to the table walk), full GPU test suite, default bench line.
set -u
OUT=gpurun_out
mkdir -p $OUT
for act in 19 12 32; do
  timeout 120 ./build/microbench_smem 32768 $act > $OUT/r2_microbench_smem_act$act.jsonl 2>&1; echo "micro act=$act exit $?"; cat $OUT/r2_microbench_smem_act$act.jsonl
done
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2_pytest_gpu_s30.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r2_pytest_gpu_s30.log
timeout 900 python bench.py > $OUT/r2_bench_default_s30.json 2> $OUT/r2_bench_default_s30.err; echo "bench exit $?"; tail -2 $OUT/r2_bench_default_s30.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_default_s30.json'))
    print('default', round(d['value'],1), d['roofline']['frac'], d['config']['kernel_variant'], d['config']['variant_ms'], d['parity']['mismatches'])
    for k,v in d['configs'].items(): print(k, round(v['value'],1), round(v['frac'],4), v['kernel'], v['parity']['mismatches'])
    print(d['next_rows']); print(d['e2e'])
except Exception as e: print('bench parse failed', e)
PY
