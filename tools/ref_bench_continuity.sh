#!/usr/bin/env bash
# The reference's own harness (tools/bench/bench.cpp via oracle/_ref/pire_bench) on this host, the way
# tools/bench/run-bench drives it: test_file doubled to >= 300 MB, whole file = ONE string.
# Single-threaded by construction (the reference has no threads).  Output: gpurun_out/ref_bench.log
OUT=gpurun_out; mkdir -p $OUT
R=oracle/_ref
F=/tmp/pire_bench_file
cp $R/test_file $F
while [ $(stat -c %s $F) -lt 300000000 ]; do cat $F $F > $F.2 && mv $F.2 $F; done
ls -la $F > $OUT/ref_bench.log
cat $F > /dev/null
for t in null nonreloc nonrelocnomask multi; do
  if [ $t = null ]; then
    echo "== -t null (memory yard-stick)" >> $OUT/ref_bench.log
    $R/pire_bench -f $F -c 3 -t null x 2>&1 | tail -3 >> $OUT/ref_bench.log
  else
    echo "== -t $t 'hello\\s+w.+d\$'" >> $OUT/ref_bench.log
    $R/pire_bench -f $F -c 3 -t $t 'hello\s+w.+d$' 2>&1 | tail -3 >> $OUT/ref_bench.log
  fi
done
echo "== -t nonreloc, the ten glued patterns" >> $OUT/ref_bench.log
$R/pire_bench -f $F -c 3 -t nonreloc 'ABCDEFGHIJKLMNOPQRSTUVWXYZ$' '[XYZ]ABCDEFGHIJKLMNOPQRSTUVWXYZ$' '[ -~]*ABCDEFGHIJKLMNOPQRSTUVWXYZ$' \
   '(\d{3}-|\(\d{3}\)\s+)(\d{3}-\d{4})$' 'hello\s+w.+d$' 'error' 'fatal' 'https?://' '^GET ' 'timeout$' 2>&1 | tail -3 >> $OUT/ref_bench.log
rm -f $F
cat $OUT/ref_bench.log
