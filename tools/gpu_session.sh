#!/usr/bin/env bash
# One gpurun call: parity tests, load-path microbenchmark, bench variants, ncu evidence.
# usage: tools/gpu_session.sh [stage ...]   (default: all)
set -u
OUT=gpurun_out
mkdir -p $OUT
STAGES="${*:-info tests micro bench ncu}"
for s in $STAGES; do
case $s in
info)
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $OUT/gpu_info.csv 2>&1
  nproc > $OUT/host_info.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host_info.txt; free -g | head -2 >> $OUT/host_info.txt
  ;;
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
  ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
  ;;
micro)
  timeout 300 ./build/microbench 4 1024 > $OUT/microbench.jsonl 2>&1; cat $OUT/microbench.jsonl
  ;;
bench)
  for wl in glue10 headline; do
    for v in plain pred priv; do
      timeout 600 python bench.py --workload $wl --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${wl}_${v}.json 2> $OUT/bench_${wl}_${v}.err
      python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${wl}_${v}.json"))
    print("$wl $v", round(d["value"],1), "GB/s frac", round(d["roofline"]["frac"],3), "clk", d["clocks"])
except Exception as e:
    print("$wl $v failed", e); print(open("$OUT/bench_${wl}_${v}.err").read()[-2000:])
PY
    done
  done
  timeout 900 python bench.py --workload utf8mixed --no-cpu > $OUT/bench_utf8mixed.json 2> $OUT/bench_utf8mixed.err; cat $OUT/bench_utf8mixed.json; tail -3 $OUT/bench_utf8mixed.err
  timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; cat $OUT/bench_default.json
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json
  ;;
ncu)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
      python bench.py --strings 2000000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant plain > $OUT/ncu_bench.log 2>&1
  for v in ${NCU_VARIANTS:-plain priv}; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniform -s 3 -c 1 -f -o $OUT/prof_glue10_$v \
        python bench.py --strings 2000000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant $v > $OUT/ncu_full_$v.log 2>&1
  done
  if [ -n "${NCU_MIXED:-}" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanGeneric -s 4 -c 1 -f -o $OUT/prof_utf8mixed \
        python bench.py --workload utf8mixed --strings 320000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant plain > $OUT/ncu_full_mixed.log 2>&1
  fi
  if [ -n "${NCU_HEADLINE:-}" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:ScanUniform -s 3 -c 1 -f -o $OUT/prof_headline_plain \
        python bench.py --workload headline --strings 2000000 --steps 2 --warmup 1 --no-e2e --no-cpu --variant plain > $OUT/ncu_full_headline.log 2>&1
  fi
  ls -la $OUT/*.ncu-rep
  ;;
esac
done
