#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for l1 in 0 1; do for v in plain pred; do
  PIRE_B200_RING_L1=$l1 timeout 300 python bench.py --workload utf8mixed --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_bench_mixed_${v}_l1$l1.json 2> $OUT/r2_bench_mixed_${v}_l1$l1.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_${v}_l1$l1.json')); print('utf8mixed $v ring_l1=$l1', round(d['value'],1), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))" || tail -3 $OUT/r2_bench_mixed_${v}_l1$l1.err
done; done
PIRE_B200_RING_L1=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "mixed or alignment or golden" > $OUT/r2_pytest_l1.log 2>&1; tail -2 $OUT/r2_pytest_l1.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > $OUT/r2_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/r2_sanitizer_memcheck.log
tail -4 $OUT/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_run.py > $OUT/r2_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $OUT/r2_sanitizer_racecheck.log
tail -4 $OUT/r2_sanitizer_racecheck.log
