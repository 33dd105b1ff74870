#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accept_sets or host_entry or default_scanner or golden" > $OUT/r2_pytest_new.log 2>&1
tail -3 $OUT/r2_pytest_new.log
timeout 900 python bench.py --no-e2e > $OUT/r2_bench_idp.json 2> $OUT/r2_bench_idp.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_idp.json"))
print("glue10", d["config"]["variant_ms"], "frac", round(d["roofline"]["frac"],4))
for k,v in d["configs"].items(): print(k, round(v["value"],1), round(v["frac"],4), v["variant_ms"], v["kernel_ms"])
print(d["next_rows"])
PY
