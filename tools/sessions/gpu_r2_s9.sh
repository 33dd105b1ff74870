#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 150 -k "uniform_bodies or mixed or alignment or golden or host_entry or accept_sets" > $OUT/r2_pytest_new.log 2>&1
tail -3 $OUT/r2_pytest_new.log
for v in plain pred look; do
  timeout 300 python bench.py --workload utf8mixed --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_bench_mixed_$v.json 2> $OUT/r2_bench_mixed_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/r2_bench_mixed_$v.json"))
    print("utf8mixed $v", round(d["value"],1), "GB/s frac", round(d["roofline"]["frac"],3), "ms", round(d["ms_per_step"],4))
except Exception as e:
    print("$v failed", e); print(open("$OUT/r2_bench_mixed_$v.err").read()[-1500:])
PY
done
timeout 300 python bench.py --workload utf8mixed --steps 5 --warmup 3 --no-e2e --no-cpu --no-configs --no-next > $OUT/r2_bench_mixed_auto.json 2> $OUT/r2_bench_mixed_auto.err
python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_auto.json')); print('mixed auto', d['config']['kernel_variant'], d['config']['variant_ms'], round(d['roofline']['frac'],3), d['parity']['mismatches'], d['parity']['checked_strings'])"
