#!/bin/bash
# Round-1 GPU pass: parity tests, bench lines, ncu launch list + full captures, the reference's own GPU backend.
mkdir -p gpurun_out
O=gpurun_out
(nproc; free -g; nvidia-smi -L; nvidia-smi topo -m) > $O/box.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench1.json 2> $O/bench1.err
timeout 300 python bench.py --impl reference > $O/bench1_ref.json 2> $O/bench1_ref.err
# launch list of the same command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_512.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/ncu_launch_run.log 2>&1
# full captures of the hot kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/dense512 \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_dense512.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_block_step -s 2 -c 1 -f -o $O/block_sphere \
    python bench.py --workload sphere --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_block.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/dense_q27f64 \
    python bench.py --workload d3q27f64 --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_q27.log 2>&1
# other configs, device-resident
for w in sphere d3q27f64 cavity64 cavity256 slab1024; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --no-e2e >> $O/bench_matrix.log 2>> $O/bench_matrix.err
done
# the UNMODIFIED reference on its own CUDA backend, same B200
for n in 128 256; do
  timeout 600 oracle/_ref/ref_lbm --device gpu --n $n --iters 60 --bench 10 --fp float --grid dGrid >> $O/ref_gpu.log 2>&1
done
timeout 600 oracle/_ref/ref_lbm --device gpu --n 256 --iters 60 --bench 10 --fp float --grid bGrid >> $O/ref_gpu.log 2>&1
timeout 900 oracle/_ref/ref_lbm --device gpu --n 512 --iters 60 --bench 10 --fp float --grid dGrid >> $O/ref_gpu.log 2>&1
timeout 300 oracle/_ref/ref_lbm --device gpu --n 48 --iters 10 --geom sphere --fp float --grid dGrid --dump $O/ref_gpu_48.bin >> $O/ref_gpu.log 2>&1
ls -la $O > $O/ls.txt
