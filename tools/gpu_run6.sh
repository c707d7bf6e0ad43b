#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/experiment.log
for e in 0 1 2 3; do
  echo "== cavity512 --experiment $e" >> $O/experiment.log
  timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu --experiment $e >> $O/experiment.log 2>&1
done
for e in 0 1 3; do
  echo "== slab1024 --experiment $e" >> $O/experiment.log
  timeout 200 python bench.py --workload slab1024 --steps 100 --warmup 10 --no-e2e --no-cpu --experiment $e >> $O/experiment.log 2>&1
  echo "== cavity256 --experiment $e" >> $O/experiment.log
  timeout 200 python bench.py --workload cavity256 --steps 200 --warmup 10 --no-e2e --no-cpu --experiment $e >> $O/experiment.log 2>&1
done
