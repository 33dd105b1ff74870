#!/usr/bin/env bash
# Session 43: pageable end-to-end path: copy threads and chunk size.
set -u
OUT=gpurun_out
mkdir -p $OUT
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
for cfg in "8 64" "12 64" "16 64" "16 32" "12 128"; do
  set -- $cfg
  PIRE_B200_HOST_THREADS=$1 PIRE_B200_HOST_CHUNK_MB=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --no-next --no-parity > $OUT/r2_e2e_t$1_c$2.json 2> $OUT/r2_e2e_t$1_c$2.err
  python -c "
import json; d=json.load(open('$OUT/r2_e2e_t$1_c$2.json')); e=d['e2e']; print('threads=$1 chunk=$2 MiB: pinned', round(e['value'],1), 'pageable', round(e['pageable']['value'],1))" || tail -3 $OUT/r2_e2e_t$1_c$2.err
done
