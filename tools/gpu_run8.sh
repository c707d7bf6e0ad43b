#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_block.py -x -q -m gpu 2>&1 | tail -5 > $O/pytest_fix.log
rm -f $O/fix.log
for w in cavity512 slab1024 cavity256 cavity128 cavity64 d3q27f64 sphere bcavity512; do
  echo "== $w" >> $O/fix.log
  timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-e2e --no-cpu >> $O/fix.log 2>&1
done
echo "== cavity512 --arith reference" >> $O/fix.log
timeout 200 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu --arith reference >> $O/fix.log 2>&1
