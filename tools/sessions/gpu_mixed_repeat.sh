#!/usr/bin/env bash
# repeat the utf8mixed bench a few times to look at step-time stability (per-step CUDA events in step_ms)
mkdir -p gpurun_out
for i in 1 2 3; do
  python bench.py --workload utf8mixed --steps 20 --no-e2e --no-cpu > gpurun_out/mixed_rep$i.json 2> gpurun_out/mixed_rep$i.err
  python - <<P
import json
j=json.loads(open("gpurun_out/mixed_rep$i.json").read().strip().splitlines()[-1])
print("rep$i", round(j["value"],1), j["ms_per_step"], j["step_ms"], j["roofline"]["kernel_ms"], j["clocks"])
P
done
PIRE_B200_NO_CLOCKS=1 python bench.py --workload utf8mixed --steps 20 --no-e2e --no-cpu > gpurun_out/mixed_rep_noclk.json 2>/dev/null
python - <<P
import json
j=json.loads(open("gpurun_out/mixed_rep_noclk.json").read().strip().splitlines()[-1])
print("noclk", round(j["value"],1), j["ms_per_step"], j["step_ms"], j["roofline"]["kernel_ms"], j["clocks"])
P
