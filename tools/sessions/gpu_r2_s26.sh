#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "split_over or mixed or utf8" > $OUT/r2_pytest_split.log 2>&1; echo "pytest split exit $?"; tail -3 $OUT/r2_pytest_split.log
for v in plain pred look; do
  timeout 300 python bench.py --workload utf8mixed --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_mixed_${v}_split.json 2> $OUT/r2_bench_mixed_${v}_split.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_${v}_split.json')); print('utf8mixed $v', round(d['value'],1), round(d['roofline']['frac'],4), round(d['ms_per_step'],4), d['parity'] and d['parity']['mismatches'])" || tail -3 $OUT/r2_bench_mixed_${v}_split.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2_launches_utf8mixed_split.csv python bench.py --workload utf8mixed --variant pred --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_ncu_launches_mixed.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_utf8mixed_split.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-6:]:
    if 'Scan' in r[4] or 'Split' in r[4]: print(r[4][:70], r[7], r[8], float(r[-1])/1e6, 'ms')
PY
