#!/usr/bin/env bash
# two GPUs: the sharded entry from C++ (world 2), bench.py under torchrun, NUMA binding, reference arm under torchrun
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2_topo_n2.txt 2>&1
timeout 600 python -m pytest tests/test_cpp_mirror.py -m gpu -x -q --timeout 300 > $OUT/r2_pytest_cpp_n2.log 2>&1; echo "pytest cpp exit $?"; tail -3 $OUT/r2_pytest_cpp_n2.log
for coll in allgather allreduce; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --collective $coll > $OUT/r2_bench_n2_$coll.json 2> $OUT/r2_bench_n2_$coll.err; echo "bench n2 $coll exit $?"
python - <<PY
import json
try:
    d=json.loads(open('$OUT/r2_bench_n2_$coll.json').read().strip().splitlines()[-1])
    print('n2 $coll', round(d['value'],1), d['ms_per_step'], d.get('parity'), d.get('e2e'))
except Exception as e:
    print('ERR', e); print(open('$OUT/r2_bench_n2_$coll.err').read()[-2000:])
PY
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-configs --no-next > $OUT/r2_bench_n1_same_box.json 2> $OUT/r2_bench_n1_same_box.err; echo "bench n1 exit $?"
python -c "
import json; d=json.loads(open('$OUT/r2_bench_n1_same_box.json').read().strip().splitlines()[-1]); print('n1', round(d['value'],1), d['ms_per_step'], d.get('e2e'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/r2_bench_ref_n2.json 2> $OUT/r2_bench_ref_n2.err; echo "ref n2 exit $?"; tail -c 600 $OUT/r2_bench_ref_n2.json
