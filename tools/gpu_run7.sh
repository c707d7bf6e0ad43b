#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/experiment2.log
for w in cavity512 slab1024; do
for r in 2 3; do
for e in 0 1; do
  echo "== $w --rpw $r --experiment $e" >> $O/experiment2.log
  timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-e2e --no-cpu --rpw $r --experiment $e >> $O/experiment2.log 2>&1
done; done; done
for r in 2 3 4; do
  echo "== cavity512 --rows-log2 $r" >> $O/experiment2.log
  timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu --rows-log2 $r >> $O/experiment2.log 2>&1
done
