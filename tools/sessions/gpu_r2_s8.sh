#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accept_sets or host_entry or default_scanner or uniform_bodies or prefix or half_final or look" > $OUT/r2_pytest_new.log 2>&1
tail -3 $OUT/r2_pytest_new.log
timeout 900 python bench.py > $OUT/r2_bench_default.json 2> $OUT/r2_bench_default.err; echo "bench exit $?"
PIRE_B200_NO_UNIFORM_BODY=1 timeout 900 python bench.py --no-e2e --no-cpu --no-parity --no-configs > $OUT/r2_bench_ringbodies.json 2> $OUT/r2_bench_ringbodies.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_default.json"))
print("glue10", d["config"]["kernel_variant"], d["config"]["variant_ms"], "frac", round(d["roofline"]["frac"],4), "e2e", d["e2e"]["value"], d["e2e"].get("pageable"))
for k,v in d["configs"].items(): print(k, round(v["value"],1), round(v["frac"],4), v["kernel_ms"])
print("uniform bodies", d["next_rows"])
print("ring bodies   ", json.load(open("gpurun_out/r2_bench_ringbodies.json"))["next_rows"])
print("parity", d["parity"]["checked_strings"], d["parity"]["mismatches"])
PY
