#!/usr/bin/env bash
# Round 2, session 1: parity of the LOOK variant + timing of every variant on the glue10 / headline corpora.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $OUT/gpu_info.csv 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "look or uniform or glued or fuzz or autoselect or full_size" > $OUT/r2_pytest_look.log 2>&1
echo "pytest exit $?" >> $OUT/r2_pytest_look.log
tail -5 $OUT/r2_pytest_look.log
for wl in glue10 headline; do
  for v in pred look look64; do
    timeout 600 python bench.py --workload $wl --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/r2_bench_${wl}_${v}.json 2> $OUT/r2_bench_${wl}_${v}.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_bench_${wl}_${v}.json"))
    print("$wl $v", round(d["value"],1), "GB/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],4), "clk", d["clocks"].get("sm_mhz"))
except Exception as e:
    print("$wl $v failed", e); print(open("$OUT/r2_bench_${wl}_${v}.err").read()[-2000:])
PY
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniformLook -s 3 -c 1 -f -o $OUT/r2_prof_glue10_look \
    python bench.py --strings 2000000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant look > $OUT/r2_ncu_look.log 2>&1
ls -la $OUT/*.ncu-rep | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniformLook -s 3 -c 1 -f -o $OUT/r2_prof_glue10_look64 \
    python bench.py --strings 2000000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant look64 > $OUT/r2_ncu_look64.log 2>&1
