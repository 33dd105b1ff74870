#!/usr/bin/env bash
# Session 39: the round's record run: full GPU test suite, smoke, default bench line, per-workload lines, reference arm.
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
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2_final_bench_reference.json 2> $OUT/r2_final_bench_reference.err; echo "reference exit $?"; head -c 600 $OUT/r2_final_bench_reference.json; echo
for wl in headline utf8mixed; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_final_bench_$wl.json 2> $OUT/r2_final_bench_$wl.err
  python -c "
import json; d=json.load(open('$OUT/r2_final_bench_$wl.json')); print('$wl', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['config']['variant_ms'], d['parity']['mismatches'])" || tail -3 $OUT/r2_final_bench_$wl.err
done
