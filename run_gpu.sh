mkdir -p gpurun_out
for i in 13 14; do ./tools/tma_probe $i; done > gpurun_out/probe.log 2>&1
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest.log
for v in "--kernel tma" "--kernel tma --workload d3q27f64" "--kernel tma --workload cavity256" "--kernel tma --workload cavity1024" "--kernel tma --arith reference"; do
  echo "== $v" >> gpurun_out/bench_var.log
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu $v 2>&1 | tail -3 >> gpurun_out/bench_var.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 1 -c 1 -o gpurun_out/prof_step_tma_r1 -f python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --kernel tma > gpurun_out/ncu_full.log 2>&1
