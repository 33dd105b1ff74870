#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2_launches_utf8mixed_split.csv python bench.py --workload utf8mixed --variant pred --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_ncu_launches_mixed.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_utf8mixed_split.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:70], r[7], r[8], float(r[-1])/1e6, 'ms')
PY
