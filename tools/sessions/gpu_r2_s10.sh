#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for alt in 0 1; do for regs in 40 48; do
  PIRE_B200_LOOK_ALT=$alt PIRE_B200_LOOK_REGS=$regs timeout 300 python bench.py --workload glue10 --variant look --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_bench_look_a${alt}_r$regs.json 2> $OUT/r2_bench_look_a${alt}_r$regs.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_look_a${alt}_r$regs.json')); print('glue10 look alt=$alt regs=$regs', round(d['value'],1), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))"
done; done
PIRE_B200_LOOK_ALT=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "look or uniform_kernel or glued" > $OUT/r2_pytest_alt.log 2>&1; tail -2 $OUT/r2_pytest_alt.log
for v in pred plain; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanGeneric -s 6 -c 1 -f -o $OUT/r2_prof_mixed_$v \
    python bench.py --workload utf8mixed --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next --variant $v > $OUT/r2_ncu_mixed_$v.log 2>&1
done
ls -la $OUT/r2_prof_mixed_*.ncu-rep
