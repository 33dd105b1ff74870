#!/usr/bin/env bash
# CSR kernel with register-streamed input (three CTAs per SM) against the cp.async ring; whole GPU suite
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/r2_pytest_gpu_full.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/r2_pytest_gpu_full.log
for ring in 0 1; do for v in plain pred look; do
  PIRE_B200_CSR_RING=$ring timeout 300 python bench.py --workload utf8mixed --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity --no-configs --no-next > $OUT/r2_bench_mixed_${v}_ring$ring.json 2> $OUT/r2_bench_mixed_${v}_ring$ring.err
  python -c "
import json; d=json.load(open('$OUT/r2_bench_mixed_${v}_ring$ring.json')); print('utf8mixed $v ring=$ring', round(d['value'],1), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))" || tail -3 $OUT/r2_bench_mixed_${v}_ring$ring.err
done; done
timeout 600 python bench.py --workload utf8mixed --steps 10 --warmup 3 > $OUT/r2_bench_utf8mixed_regs.json 2> $OUT/r2_bench_utf8mixed_regs.err; tail -c 1500 $OUT/r2_bench_utf8mixed_regs.json
