#!/bin/bash
# four GPUs: bench.py at N=4 for every multi-GPU config of BASELINE.json, and the C++ app driving 4 GPUs from one process
mkdir -p gpurun_out
O=gpurun_out
APP=neon_b200/cpp/bin/lbm-lid-driven-cavity-flow
(nproc; free -g; nvidia-smi -L; nvidia-smi topo -m) > $O/box4.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 4 > $O/bench4.json 2> $O/bench4.err
timeout 300 $TR --master-port 29522 bench.py --gpus 4 --workload cavity1024 --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench4_strong.json 2> $O/bench4_strong.err
timeout 300 $TR --master-port 29523 bench.py --gpus 4 --workload sphere --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench4_sphere.json 2> $O/bench4_sphere.err
timeout 300 $TR --master-port 29524 bench.py --gpus 4 --workload d3q27f64 --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench4_q27.json 2> $O/bench4_q27.err
timeout 300 $TR --master-port 29525 bench.py --gpus 4 --transport fused --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench4_fused.json 2> $O/bench4_fused.err
cd $O
B="--computeFP float --storageFP float --benchmark --warmup-iter 10 --max-iter 110 --device-setup"
( echo "== dGrid 1024x1024x512 4 GPUs --sOCC --put"
  timeout 300 ../$APP --deviceType gpu --deviceIds 0 1 2 3 --grid dGrid --dim 1024 1024 512 --sOCC --put $B --report-filename cpp4
  echo "== dGrid 1024x1024x512 4 GPUs --nOCC --get"
  timeout 300 ../$APP --deviceType gpu --deviceIds 0 1 2 3 --grid dGrid --dim 1024 1024 512 --nOCC --get $B --report-filename cpp4
  echo "== bGrid 1024x512x512 4 GPUs --sOCC --put"
  timeout 300 ../$APP --deviceType gpu --deviceIds 0 1 2 3 --grid bGrid --dim 1024 512 512 --sOCC --put $B --report-filename cpp4
) > cpp_app4.log 2>&1
grep -h "^==\|MLUPS:\|Problem Setup\|Grid Init\|Exception" cpp_app4.log > cpp_app4_metrics.log
rm -f cpp4_*.json
