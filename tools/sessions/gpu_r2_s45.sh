#!/usr/bin/env bash
# Session 45: record run of the code as it ends the round (host-entry capacity fix, LOOKH removed again):
# full GPU test suite, smoke, default bench line, reference arm.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2_final_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2_final_pytest_gpu.log; tail -3 $OUT/r2_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/r2_final_smoke.log
SECONDS=0
timeout 900 python bench.py > $OUT/r2_final_bench_default.json 2> $OUT/r2_final_bench_default.err; echo "bench exit $? after ${SECONDS}s"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2_final_bench_default.json'))
    print('value', round(d['value'], 1), 'frac', round(d['roofline']['frac'], 4), d['roofline']['kernel'], 'parity', d['parity']['mismatches'], '/', d['parity']['checked_strings'])
    print('e2e', round(d['e2e']['value'], 1), 'pageable', round(d['e2e']['pageable']['value'], 1), 'cpu', round(d['cpu_baseline']['value'], 2), 'clocks', d['clocks'])
    for k, v in d['configs'].items():
        print(k, round(v['value'], 1), round(v['roofline']['frac'], 4), 'parity', v['parity']['mismatches'])
    print('next_rows', {k: round(v['value'], 1) for k, v in d['next_rows'].items() if isinstance(v, dict)})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2_final_bench_default.err').read()[-2000:])
PY
SECONDS=0
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2_final_bench_reference.json 2> $OUT/r2_final_bench_reference.err; echo "reference exit $? after ${SECONDS}s"; head -c 400 $OUT/r2_final_bench_reference.json; echo
