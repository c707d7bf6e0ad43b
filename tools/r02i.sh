#!/bin/bash
# round 2, GPU call I: does the distance between the 19 population planes (pitch_q = planes x 1 MB at 512 x 512) matter?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --no-e2e --no-cpu --no-extras --steps 100 --warmup 10"
for nz in 512 513 514 516 520 528 544 576 640; do
  timeout 300 $B --workload box512x512x$nz > $O/r02i_pitch_$nz.json 2> $O/r02i_pitch_$nz.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02i_pitch_$nz.json").read().strip().splitlines()[-1])
    print("512x512x$nz", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4))
except Exception as e:
    print("$nz FAILED", e)
PY
done 2>&1 | tee $O/r02i_pitch_sweep.log
for nz in 128 129 130 132 136 144; do
  timeout 300 $B --workload box1024x1024x$nz > $O/r02i_pitchs_$nz.json 2> $O/r02i_pitchs_$nz.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02i_pitchs_$nz.json").read().strip().splitlines()[-1])
    print("1024x1024x$nz", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4))
except Exception as e:
    print("$nz FAILED", e)
PY
done 2>&1 | tee -a $O/r02i_pitch_sweep.log
# small boxes: cells per thread (more, shorter threads) and block shape, with 10 iterations per CUDA-graph replay
for w in cavity64 cavity96 cavity128; do
  for v in "--vec 4" "--vec 2" "--vec 1" "--vec 2 --rows-log2 4" "--vec 1 --rows-log2 4" "--vec 2 --rows-log2 3" "--vec 4 --graph-iters 50" "--vec 2 --graph-iters 50" "--vec 4 --graph-iters 0"; do
    timeout 300 $B --workload $w $v > $O/r02i_small.json 2> $O/r02i_small.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02i_small.json").read().strip().splitlines()[-1])
    print("$w $v:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step")
except Exception as e:
    print("$w $v FAILED", e)
PY
  done
done 2>&1 | tee $O/r02i_small_sweep.log
