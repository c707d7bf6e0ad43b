#!/bin/bash
# Multi-GPU pass on one box: bench.py at N GPUs (one process per GPU) for every multi-GPU config of BASELINE.json, then the
# C++ benchmark app driving the same N GPUs from ONE process.   gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
APP=neon_b200/cpp/bin/lbm-lid-driven-cavity-flow
(nproc; free -g; nvidia-smi -L; nvidia-smi topo -m) > $O/box$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N > $O/bench$N.json 2> $O/bench$N.err
timeout 300 $TR --master-port 29522 bench.py --gpus $N --workload cavity1024 --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench${N}_strong.json 2> $O/bench${N}_strong.err
timeout 300 $TR --master-port 29523 bench.py --gpus $N --workload sphere --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench${N}_sphere.json 2> $O/bench${N}_sphere.err
timeout 300 $TR --master-port 29524 bench.py --gpus $N --workload d3q27f64 --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench${N}_q27.json 2> $O/bench${N}_q27.err
timeout 300 $TR --master-port 29525 bench.py --gpus $N --transport fused --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench${N}_fused.json 2> $O/bench${N}_fused.err
IDS=$(seq -s ' ' 0 $((N-1)))
cd $O
B="--computeFP float --storageFP float --benchmark --warmup-iter 10 --max-iter 110 --device-setup"
( echo "== dGrid 1024x1024x$((128*N)) $N GPUs --sOCC --put"
  timeout 300 ../$APP --deviceType gpu --deviceIds $IDS --grid dGrid --dim 1024 1024 $((128*N)) --sOCC --put $B --report-filename cppN
  echo "== dGrid 1024x1024x$((128*N)) $N GPUs --nOCC --get"
  timeout 300 ../$APP --deviceType gpu --deviceIds $IDS --grid dGrid --dim 1024 1024 $((128*N)) --nOCC --get $B --report-filename cppN
  echo "== bGrid 1024x512x512 $N GPUs --sOCC --put"
  timeout 300 ../$APP --deviceType gpu --deviceIds $IDS --grid bGrid --dim 1024 512 512 --sOCC --put $B --report-filename cppN
) > cpp_app$N.log 2>&1
grep -h "^==\|MLUPS:\|Problem Setup\|Grid Init\|Exception" cpp_app$N.log > cpp_app${N}_metrics.log
rm -f cppN_*.json
