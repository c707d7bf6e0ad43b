#!/bin/bash
# 8 x B200 (charged 8x: kept short): the default bench line with every extra config, the device-side timeline of one
# iteration, A/B of OCC, the C++ app from one process
N=${1:-8}
T=${2:-r02h}
mkdir -p gpurun_out
O=gpurun_out
APP=neon_b200/cpp/bin/lbm-lid-driven-cavity-flow
(nproc; free -g; nvidia-smi -L; nvidia-smi topo -m) > $O/${T}_box$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $N > $O/${T}_bench$N.json 2> $O/${T}_bench$N.err
Q="--steps 100 --warmup 10 --no-e2e --no-cpu --no-extras"
for v in "nocc:--occ none" "nopipe:--no-pipeline"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 $TR --master-port 29522 bench.py --gpus $N $Q $flags > $O/${T}_bench${N}_$name.json 2> $O/${T}_bench${N}_$name.err
done
timeout 300 $TR --master-port 29523 tools/halo_timeline.py > $O/${T}_timeline$N.txt 2> $O/${T}_timeline$N.err
IDS=$(seq -s ' ' 0 $((N-1)))
cd $O
B="--computeFP float --storageFP float --benchmark --warmup-iter 10 --max-iter 110 --device-setup"
( echo "== dGrid 1024x1024x$((128*N)) $N GPUs --sOCC --put"
  timeout 300 ../$APP --deviceType gpu --deviceIds $IDS --grid dGrid --dim 1024 1024 $((128*N)) --sOCC --put $B --report-filename cppN
) > ${T}_cpp_app$N.log 2>&1
grep -h "^==\|MLUPS:\|Problem Setup\|Grid Init\|Exception" ${T}_cpp_app$N.log > ${T}_cpp_app${N}_metrics.log
rm -f cppN_*.json
