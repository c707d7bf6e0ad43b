#!/bin/bash
# Multi-GPU pass on one box: bench.py at N GPUs (one process per GPU; the default line carries every multi-GPU config of
# BASELINE.json as extra_configs), A/B of the halo schedule, the device-side timeline of one iteration, the multi-GPU tests,
# then the C++ benchmark app driving the same N GPUs from ONE process.   gpurun --gpus N -- 'bash tools/gpu_multi.sh N TAG'
N=${1:-2}
T=${2:-r02}
mkdir -p gpurun_out
O=gpurun_out
APP=neon_b200/cpp/bin/lbm-lid-driven-cavity-flow
(nproc; free -g; nvidia-smi -L; nvidia-smi topo -m) > $O/${T}_box$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus $N > $O/${T}_bench$N.json 2> $O/${T}_bench$N.err
Q="--steps 100 --warmup 10 --no-e2e --no-cpu --no-extras"
for v in "pipelined:" "nopipe:--no-pipeline" "nocc:--occ none" "nocc_nopipe:--occ none --no-pipeline" "fused:--transport fused" "packed:--transport packed"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 $TR --master-port 29522 bench.py --gpus $N $Q $flags > $O/${T}_bench${N}_$name.json 2> $O/${T}_bench${N}_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${T}_bench${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), round(j["ms_per_step"],4), "kernel alone", round(j["roofline"]["kernel_ms"],4), round(j["roofline"]["frac"],4))
except Exception as e:
    print("$name FAILED", e)
PY
done
timeout 300 $TR --master-port 29523 tools/halo_timeline.py > $O/${T}_timeline$N.txt 2> $O/${T}_timeline$N.err
timeout 300 $TR --master-port 29524 tools/halo_timeline.py --no-pipeline > $O/${T}_timeline${N}_nopipe.txt 2>> $O/${T}_timeline$N.err
timeout 300 $TR --master-port 29525 tools/halo_timeline.py --occ none > $O/${T}_timeline${N}_nocc.txt 2>> $O/${T}_timeline$N.err
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multiproc.py tests/test_cpp_veneer.py -q -m gpu > $O/${T}_pytest_multi$N.log 2>&1
  tail -4 $O/${T}_pytest_multi$N.log
fi
IDS=$(seq -s ' ' 0 $((N-1)))
cd $O
B="--computeFP float --storageFP float --benchmark --warmup-iter 10 --max-iter 110 --device-setup"
( echo "== dGrid 1024x1024x$((128*N)) $N GPUs --sOCC --put"
  timeout 300 ../$APP --deviceType gpu --deviceIds $IDS --grid dGrid --dim 1024 1024 $((128*N)) --sOCC --put $B --report-filename cppN
  echo "== dGrid 1024x1024x$((128*N)) $N GPUs --nOCC --get"
  timeout 300 ../$APP --deviceType gpu --deviceIds $IDS --grid dGrid --dim 1024 1024 $((128*N)) --nOCC --get $B --report-filename cppN
  echo "== bGrid 1024x512x512 $N GPUs --sOCC --put"
  timeout 300 ../$APP --deviceType gpu --deviceIds $IDS --grid bGrid --dim 1024 512 512 --sOCC --put $B --report-filename cppN
) > ${T}_cpp_app$N.log 2>&1
grep -h "^==\|MLUPS:\|Problem Setup\|Grid Init\|Exception" ${T}_cpp_app$N.log > ${T}_cpp_app${N}_metrics.log
rm -f cppN_*.json
