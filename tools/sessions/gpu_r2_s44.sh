#!/usr/bin/env bash
# Session 44: LOOKH variant (look-ahead filter with hashed slots): parity in both modes (even bytes hashed = default,
# every byte hashed), then the glued scan with each mode and with LOOK on the same box.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lookh or dense_near or auto_picks or host_entry" > $OUT/r2_pytest_s44a.log 2>&1; echo "pytest (even bytes hashed) exit $?"; tail -3 $OUT/r2_pytest_s44a.log
PIRE_B200_LOOKH_MODE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lookh or dense_near" > $OUT/r2_pytest_s44b.log 2>&1; echo "pytest (every byte hashed) exit $?"; tail -3 $OUT/r2_pytest_s44b.log
show() {
  python -c "
import json,sys; d=json.load(open('$1')); r=d['roofline']; print('$2', round(d['value'],1), 'GB/s', round(r['kernel_ms'],4), 'ms', round(r['frac'],4), r['kernel'], 'parity', d.get('parity',{}).get('mismatches'), d.get('parity',{}).get('checked_strings'))" || tail -5 ${1%.json}.err
}
for mode in 2 1; do
  PIRE_B200_LOOKH_MODE=$mode timeout 400 python bench.py --workload glue10 --variant lookh --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_glue10_lookh_m$mode.json 2> $OUT/r2_bench_glue10_lookh_m$mode.err
  show $OUT/r2_bench_glue10_lookh_m$mode.json "lookh mode $mode:"
done
timeout 400 python bench.py --workload glue10 --variant look --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next --no-parity > $OUT/r2_bench_glue10_look_s44.json 2> $OUT/r2_bench_glue10_look_s44.err
show $OUT/r2_bench_glue10_look_s44.json "look (same box):"
timeout 400 python bench.py --workload headline --variant lookh --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_headline_lookh.json 2> $OUT/r2_bench_headline_lookh.err
show $OUT/r2_bench_headline_lookh.json "headline lookh:"
