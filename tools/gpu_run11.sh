#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/experiment4.log
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_cpp_veneer.py -x -q -m gpu 2>&1 | tail -5 > $O/pytest_fix.log
for w in cavity512 slab1024 cavity256 cavity128 cavity64 d3q27f64 cavity1024; do
  echo "== $w" >> $O/experiment4.log
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-e2e --no-cpu >> $O/experiment4.log 2>&1
done
echo "== cavity512 reference" >> $O/experiment4.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu --arith reference >> $O/experiment4.log 2>&1
