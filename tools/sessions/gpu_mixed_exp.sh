#!/usr/bin/env bash
# experiments on the ragged (utf8mixed) workload: resident CTAs per SM
OUT=gpurun_out; mkdir -p $OUT
for c in 1 2 3; do
  PIRE_B200_GENERIC_CTAS=$c timeout 600 python bench.py --workload utf8mixed --steps 5 --warmup 3 --no-e2e --no-cpu --variant plain > $OUT/mixed_ctas$c.json 2> $OUT/mixed_ctas$c.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/mixed_ctas$c.json")); print("ctas $c:", round(d["value"],1), "GB/s", round(d["ms_per_step"],2), "ms")
except Exception as e:
    print("ctas $c failed", e); print(open("$OUT/mixed_ctas$c.err").read()[-1500:])
PY
done
