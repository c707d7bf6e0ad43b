mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for v in "--vec 4" "--vec 2" "--vec 1" "--arith reference" "--rows-log2 1" "--rows-log2 2" "--rows-log2 3" "--rows-log2 4"; do
  echo "== $v" >> gpurun_out/bench_var.log
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu $v >> gpurun_out/bench_var.log 2>&1
done
for w in cavity64 cavity128 cavity256 d3q27f64 cavity1024; do
  echo "== $w" >> gpurun_out/bench_var.log
  python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --workload $w >> gpurun_out/bench_var.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 1 -c 2 -o gpurun_out/prof_step_r1 -f python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
nvidia-smi > gpurun_out/smi.txt
