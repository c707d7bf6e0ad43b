#!/bin/bash
# round 2, GPU call V: the 64^3 step kernel with a WARM L2 under ncu (--cache-control none), both cell-per-thread widths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for v in 2 4; do
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_dense_step -s 60 -c 1 -f -o $O/r02v_step64_warm_v$v \
    python bench.py --workload cavity64 --graph-iters 0 --vec $v --steps 60 --warmup 20 --no-cpu --no-e2e --no-extras > $O/r02v_ncu_v$v.log 2>&1
done
B="timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 400 --warmup 40"
for w in cavity64 cavity96 cavity128 cavity192; do
  for v in "" "--graph-iters 10" "--graph-iters 100"; do
    $B --workload $w $v > $O/r02v_small.json 2> $O/r02v_small.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02v_small.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step", "frac", round(j["roofline"]["frac"],3), j["config"]["issue"], j["config"]["graph_iters"])
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02v_small.err").read()[-400:])
PY
  done
done 2>&1 | tee $O/r02v_small_sweep.log
