#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest.log
: > gpurun_out/bench_var.log
for v in "" "--flags-summary-first" "--workload d3q27f64" "--workload d3q27f64 --vec 2" "--workload d3q27f64 --kernel tma" "--workload d3q27f64 --flags-summary-first" \
   "--workload cavity64" "--workload cavity128" "--workload cavity256" "--arith reference" "--workload slab1024"; do
  echo "== $v" >> gpurun_out/bench_var.log
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu $v 2>&1 | tail -3 >> gpurun_out/bench_var.log
done
