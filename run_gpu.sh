#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest.log
: > gpurun_out/bench_var.log
for v in "--kernel direct" "--kernel direct --flags-always" "--kernel direct --flags-always --workload cavity1024" "--kernel direct --flags-always --workload cavity256" \
         "--kernel direct --flags-always --workload d3q27f64" "--kernel direct --flags-always --vec 2" "--kernel direct --flags-always --arith reference"; do
  echo "== $v" >> gpurun_out/bench_var.log
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu $v 2>&1 | tail -3 >> gpurun_out/bench_var.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -o gpurun_out/prof_step_direct_r1f -f \
  python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --kernel direct --flags-always > gpurun_out/ncu_full.log 2>&1
