#!/usr/bin/env bash
# Session 49: what does sampling the clocks cost the timed region?  NVML + nvidia-smi (as before), NVML alone at 4 / 10 ms, none.
set -u
OUT=gpurun_out
mkdir -p $OUT
run() {
  env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-configs --no-next --no-parity --variant look > $OUT/r2_clk.json 2> $OUT/r2_clk.err
  python -c "
import json; d=json.load(open('$OUT/r2_clk.json')); print('$*', 'ms/step', round(d['ms_per_step'],4), 'median', round(d['step_ms']['median'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'samples', (d.get('clocks') or {}).get('samples'))" || tail -3 $OUT/r2_clk.err
}
run PIRE_B200_CLOCKS_SMI=1
run PIRE_B200_CLOCKS_POLL_MS=4
run PIRE_B200_CLOCKS_POLL_MS=10
run PIRE_B200_NO_CLOCKS=1
run PIRE_B200_CLOCKS_SMI=1
run PIRE_B200_CLOCKS_POLL_MS=10
