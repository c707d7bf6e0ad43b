#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest.log
: > gpurun_out/bench_var.log
for v in "" "--workload cavity256" "--workload cavity256 --rpw 1" "--workload cavity256 --rpw 2" "--workload cavity256 --rpw 3" "--workload cavity128" "--workload cavity128 --rpw 1" "--workload cavity128 --rpw 2" "--workload cavity128 --rpw 3" \
   "--workload cavity64" "--workload cavity64 --rpw 1" "--workload cavity64 --rpw 3" "--workload cavity1024" "--workload d3q27f64" "--workload slab1024"; do
  echo "== $v" >> gpurun_out/bench_var.log
  timeout 120 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu $v 2>&1 | tail -3 >> gpurun_out/bench_var.log
done
