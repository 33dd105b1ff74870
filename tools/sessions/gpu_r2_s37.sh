#!/usr/bin/env bash
# Session 37: evidence for the kernels that ship (full ncu captures, launch list), split-kernel profile and threshold sweep,
# sanitizer passes over every kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
# 1. full ncu capture of the default glued scan (look-ahead, two strings per lane) at bench size
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniformLook2 -s 4 -c 1 -f -o $OUT/r2_prof_full_glue10_look2 \
    python bench.py --workload glue10 --variant look --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-next --no-parity > $OUT/r2_ncu_full_look2.log 2>&1
# 2. launch list of the default bench command (short)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/r2_launches_default.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > $OUT/r2_ncu_launches_default.log 2>&1
# 3. split kernel profile on the mixed corpus
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanSplit -s 3 -c 1 -f -o $OUT/r2_prof_full_utf8mixed_split \
    python bench.py --workload utf8mixed --variant pred --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --no-next --no-parity > $OUT/r2_ncu_full_split.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2_launches_utf8mixed.csv \
    python bench.py --workload utf8mixed --variant pred --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_ncu_launches_mixed.log 2>&1
# 4. split threshold sweep
for m in 4096 8192 16384; do
  PIRE_B200_SPLIT_MIN=$m timeout 300 python bench.py --workload utf8mixed --variant pred --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_mixed_split$m.json 2> $OUT/r2_bench_mixed_split$m.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_split$m.json')); print('utf8mixed split_min=$m', round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],4), d['parity']['mismatches'])" || tail -3 $OUT/r2_bench_mixed_split$m.err
done
# 5. sanitizer
bash tools/gpu_sanitize.sh
ls -la $OUT/*.ncu-rep
