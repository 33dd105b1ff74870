#!/usr/bin/env bash
# Session 48 (two GPUs): the sharded bench line of the code as the round ends, launched the way the driver launches it.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2_final_bench_n2.json 2> $OUT/r2_final_bench_n2.err; echo "bench N=2 exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_final_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'parity', d['parity'], 'e2e', {k: v for k, v in d['e2e'].items() if k in ('value', 'numa')}, 'pageable', d['e2e'].get('pageable'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r2_final_bench_n2.err').read()[-2500:])
PY
