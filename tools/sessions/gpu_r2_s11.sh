#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for ctas in 2 3; do for v in plain pred look; do
  PIRE_B200_GENERIC_CTAS=$ctas timeout 300 python bench.py --workload utf8mixed --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_bench_mixed_${v}_c$ctas.json 2> $OUT/r2_bench_mixed_${v}_c$ctas.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_${v}_c$ctas.json')); print('utf8mixed $v ctas=$ctas', round(d['value'],1), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))" || tail -3 $OUT/r2_bench_mixed_${v}_c$ctas.err
done; done
# evidence at bench size for the glued scan: launch list + full captures of the exit-filter and look-ahead kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r2_launches_glue10_10GB.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_ncu_launches_glue10.log 2>&1
for v in pred look; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniform -s 6 -c 1 -f -o $OUT/r2_prof_full_glue10_$v \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next --variant $v > $OUT/r2_ncu_full_glue10_$v.log 2>&1
done
ls -la $OUT/r2_prof_full_*.ncu-rep
