#!/usr/bin/env bash
# multi-GPU session: bench.py under torchrun at N = $1 (default 2)
set -u
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$N.csv 2>&1
for wl in ${WORKLOADS:-glue10}; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --workload $wl > $OUT/bench_${wl}_n$N.json 2> $OUT/bench_${wl}_n$N.err
  tail -1 $OUT/bench_${wl}_n$N.json; tail -3 $OUT/bench_${wl}_n$N.err
done
