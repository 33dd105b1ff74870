#!/usr/bin/env bash
# Round 2: LOOK register budgets (40 regs x 512 threads x 3 CTAs vs 48 regs x 384 threads x 3 CTAs)
set -u
OUT=gpurun_out
mkdir -p $OUT
for regs in 48 40; do
  for v in look look64; do
    PIRE_B200_LOOK_REGS=$regs timeout 600 python bench.py --workload glue10 --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/r2_bench_glue10_${v}_r$regs.json 2> $OUT/r2_bench_glue10_${v}_r$regs.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_bench_glue10_${v}_r$regs.json"))
    print("glue10 $v regs=$regs", round(d["value"],1), "GB/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],4), "clk", d["clocks"].get("sm_mhz"))
except Exception as e:
    print("$v failed", e); print(open("$OUT/r2_bench_glue10_${v}_r$regs.err").read()[-2000:])
PY
  done
done
PIRE_B200_LOOK_REGS=${NCU_REGS:-48} timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniformLook -s 3 -c 1 -f -o $OUT/r2_prof_glue10_look \
    python bench.py --strings 2000000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant look > $OUT/r2_ncu_look.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "look or uniform or glued" > $OUT/r2_pytest_look.log 2>&1
tail -3 $OUT/r2_pytest_look.log
