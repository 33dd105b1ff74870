#!/usr/bin/env bash
# Session 38: split kernel with an L2 prefetch ahead of every lane's walk.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_over or mixed or utf8" > $OUT/r2_pytest_s38.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2_pytest_s38.log
for pf in 16 0 8 32 64; do
  PIRE_B200_SPLIT_PREFETCH=$pf timeout 300 python bench.py --workload utf8mixed --variant pred --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_mixed_pf$pf.json 2> $OUT/r2_bench_mixed_pf$pf.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_pf$pf.json')); print('utf8mixed prefetch=$pf', round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],4), d['parity']['mismatches'])" || tail -3 $OUT/r2_bench_mixed_pf$pf.err
done
