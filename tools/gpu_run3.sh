#!/bin/bash
# two GPUs: the C++ veneer's single-process multi-GPU path, and bench.py at N=2 (one process per GPU) with the new e2e leg
mkdir -p gpurun_out
O=gpurun_out
APP=neon_b200/cpp/bin/lbm-lid-driven-cavity-flow
(nproc; free -g; nvidia-smi -L; nvidia-smi topo -m) > $O/box2.txt 2>&1
timeout 900 python -m pytest tests/test_cpp_veneer.py -x -q -m gpu -k "bgrid or two_real" 2>&1 | tail -25 > $O/pytest_cpp2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 > $O/bench2.json 2> $O/bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --impl reference > $O/bench2_ref.json 2> $O/bench2_ref.err
cd $O
B="--computeFP float --storageFP float --benchmark --warmup-iter 10 --max-iter 110 --device-setup"
( for opt in "--sOCC --get" "--sOCC --put" "--nOCC --get" "--nOCC --put" "--sOCC --put --huGrid"; do
    echo "== dGrid 1024x1024x256 2 GPUs $opt"
    timeout 300 ../$APP --deviceType gpu --deviceIds 0 1 --grid dGrid --dim 1024 1024 256 $opt $B --report-filename cpp2_dgrid
  done
  echo "== bGrid 1024x512x512 2 GPUs --sOCC --put"
  timeout 600 ../$APP --deviceType gpu --deviceIds 0 1 --grid bGrid --dim 1024 512 512 --sOCC --put $B --report-filename cpp2_bgrid
  echo "== bGrid 512^3 1 GPU"
  timeout 600 ../$APP --deviceType gpu --deviceIds 0 --grid bGrid --domain-size 512 $B --report-filename cpp1_bgrid
  echo "== dGrid 512^3 1 GPU (grid init without host mirrors)"
  timeout 600 ../$APP --deviceType gpu --deviceIds 0 --grid dGrid --domain-size 512 $B --report-filename cpp1_dgrid
) > cpp_app2.log 2>&1
grep -h "^==\|MLUPS:\|Problem Setup\|Grid Init\|Exception" cpp_app2.log > cpp_app2_metrics.log
rm -f cpp2_*.json cpp1_*.json
