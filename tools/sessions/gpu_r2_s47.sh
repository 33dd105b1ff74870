#!/usr/bin/env bash
# Session 47: staging copy of pageable input with non-temporal stores: host-entry tests, then the end-to-end numbers three times.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host" > $OUT/r2_pytest_s47.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2_pytest_s47.log
for rep in 1 2 3; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --no-next --no-parity > $OUT/r2_e2e_nt_$rep.json 2> $OUT/r2_e2e_nt_$rep.err
  python -c "
import json; d=json.load(open('$OUT/r2_e2e_nt_$rep.json')); e=d['e2e']; print('run $rep: pinned', round(e['value'],1), 'pageable', round(e['pageable']['value'],1))" || tail -3 $OUT/r2_e2e_nt_$rep.err
done
PIRE_B200_HOST_THREADS=14 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --no-next --no-parity > $OUT/r2_e2e_nt_t14.json 2> $OUT/r2_e2e_nt_t14.err
python -c "
import json; d=json.load(open('$OUT/r2_e2e_nt_t14.json')); e=d['e2e']; print('14 threads: pinned', round(e['value'],1), 'pageable', round(e['pageable']['value'],1))"
