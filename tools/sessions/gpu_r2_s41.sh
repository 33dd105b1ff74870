#!/usr/bin/env bash
# Session 41: record run of the final code (tests, smoke, default bench, reference arm) + sanitizer passes.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2_final_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2_final_pytest_gpu.log; tail -3 $OUT/r2_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/r2_final_smoke.log
timeout 900 python bench.py > $OUT/r2_final_bench_default.json 2> $OUT/r2_final_bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_final_bench_default.json'))
    print('default', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['config']['variant_ms'], 'mismatches', d['parity']['mismatches'], 'clocks', d['clocks'])
    for k,v in d['configs'].items(): print(k, round(v['value'],1), round(v['frac'],4), v['kernel'], v['parity']['mismatches'])
    print(d['next_rows']); print({k:d['e2e'][k] for k in ('value','pageable')}); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r2_final_bench_default.err').read()[-2000:])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2_final_bench_reference.json 2> $OUT/r2_final_bench_reference.err; echo "reference exit $?"
bash tools/gpu_sanitize.sh
