#!/usr/bin/env bash
# long strings split over a warp: parity, then the mixed-length bench with and without it
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "split_over or mixed or utf8 or fuzz or host_entry" > $OUT/r2_pytest_split.log 2>&1; echo "pytest split exit $?"; tail -8 $OUT/r2_pytest_split.log
for sp in 1 0; do for v in plain pred; do
  PIRE_B200_SPLIT=$sp timeout 300 python bench.py --workload utf8mixed --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_mixed_${v}_split$sp.json 2> $OUT/r2_bench_mixed_${v}_split$sp.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_${v}_split$sp.json')); print('utf8mixed $v split=$sp', round(d['value'],1), round(d['roofline']['frac'],4), round(d['ms_per_step'],4), d['parity'] and d['parity']['mismatches'])" || tail -3 $OUT/r2_bench_mixed_${v}_split$sp.err
done; done
