#!/bin/bash
# one-GPU pass: full GPU suite, bench line, launch list, ncu --set full captures, bench matrix
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/pytest_gpu_final.log
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_512_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/dense512_final \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_dense512.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_block_step -s 2 -c 1 -f -o $O/block_sphere_final \
    python bench.py --workload sphere --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_block.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/dense_q27f64_final \
    python bench.py --workload d3q27f64 --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_q27.log 2>&1
rm -f $O/bench_matrix_final.log
for w in sphere bcavity512 d3q27f64 cavity64 cavity128 cavity256 slab1024 cavity1024; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --no-e2e >> $O/bench_matrix_final.log 2>> $O/bench_matrix_final.err
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --arith reference >> $O/bench_matrix_final.log 2>> $O/bench_matrix_final.err
for e in 3; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --experiment $e >> $O/bench_matrix_final.log 2>> $O/bench_matrix_final.err
done
