#!/usr/bin/env bash
# Session 40: LOOK1 variant id, split kernel with the head of every piece asked for before the stitch.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2_pytest_s40.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r2_pytest_s40.log
for i in 1 2; do
timeout 300 python bench.py --workload utf8mixed --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_mixed_head$i.json 2> $OUT/r2_bench_mixed_head$i.err
python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_head$i.json')); print('utf8mixed head-preload', round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],4), d['config']['variant_ms'], d['parity']['mismatches'])" || tail -3 $OUT/r2_bench_mixed_head$i.err
done
timeout 300 python bench.py --workload glue10 --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_glue10_s40.json 2> $OUT/r2_bench_glue10_s40.err
python -c "
import json; d=json.load(open('$OUT/r2_bench_glue10_s40.json')); print('glue10', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['config']['variant_ms'], d['parity']['mismatches'])" || tail -3 $OUT/r2_bench_glue10_s40.err
