#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_block.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_block.log
: > gpurun_out/bench_var.log
for v in "--workload bcavity256" "--workload bcavity512" "--workload sphere"; do
  echo "== $v" >> gpurun_out/bench_var.log
  timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu $v 2>&1 | tail -3 >> gpurun_out/bench_var.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_block_step -s 2 -c 1 -o gpurun_out/prof_block_r1b -f \
  python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --workload bcavity512 > gpurun_out/ncu_full.log 2>&1
