#!/usr/bin/env bash
# Session 32: look-ahead kernel (clean bits): CTA shapes, then one full ncu capture of the default shape.
set -u
OUT=gpurun_out
mkdir -p $OUT
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload glue10 --variant look --steps 20 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_$name.json 2> $OUT/r2_bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('$OUT/r2_bench_$name.json')); print('$name', round(d['value'],1), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms'],4), d['parity']['mismatches'])
except Exception as e: print('$name failed', e); print(open('$OUT/r2_bench_$name.err').read()[-1500:])
PY
}
run look_b640 PIRE_B200_LOOK_BLOCK=640
run look_b320 PIRE_B200_LOOK_BLOCK=320
run look_b256 PIRE_B200_LOOK_BLOCK=256
run look_b384 PIRE_B200_LOOK_BLOCK=384
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniformLook -s 6 -c 1 -f -o $OUT/r2_prof_full_glue10_lookclean \
    python bench.py --workload glue10 --variant look --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --no-next --no-parity > $OUT/r2_ncu_full_lookclean.log 2>&1
ls -la $OUT/*.ncu-rep
