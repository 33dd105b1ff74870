#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "split_over or mixed or utf8" > $OUT/r2_pytest_split.log 2>&1; echo "pytest split exit $?"; tail -3 $OUT/r2_pytest_split.log
for m in 4096 8192 16384; do for v in pred; do
  PIRE_B200_SPLIT_MIN=$m timeout 300 python bench.py --workload utf8mixed --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_mixed_${v}_min$m.json 2> $OUT/r2_bench_mixed_${v}_min$m.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_${v}_min$m.json')); print('utf8mixed $v split_min=$m', round(d['value'],1), round(d['roofline']['frac'],4), round(d['ms_per_step'],4), d['parity'] and d['parity']['mismatches'])" || tail -3 $OUT/r2_bench_mixed_${v}_min$m.err
done; done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ScanSplitKernel -s 3 -c 1 -o $OUT/r2_prof_split_kernel -f python bench.py --workload utf8mixed --variant pred --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_ncu_split.log 2>&1; tail -2 $OUT/r2_ncu_split.log
