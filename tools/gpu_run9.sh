#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/experiment3.log
timeout 600 python -m pytest tests/test_gpu_dense.py -x -q -m gpu 2>&1 | tail -3 > $O/pytest_fix.log
for w in cavity512 slab1024 cavity256 cavity128 d3q27f64; do
  echo "== $w" >> $O/experiment3.log
  timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-e2e --no-cpu >> $O/experiment3.log 2>&1
done
