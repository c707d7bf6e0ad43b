#!/bin/bash
# round 2, GPU call N: 4 blocks/SM (64 registers) for two cells per thread: one wave at 64^3?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 400 --warmup 40"
for w in cavity48 cavity64 cavity80 cavity96 cavity128; do
  for v in "" "--vec 4" "--persistent --chain-graph" "--persistent --chain-graph --vec 2"; do
    $B --workload $w --graph-iters 10 $v > $O/r02n_small.json 2> $O/r02n_small.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02n_small.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step", "frac", round(j["roofline"]["frac"],3), "iters/launch", j["config"]["iterations_per_launch"])
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02n_small.err").read()[-400:])
PY
  done
done 2>&1 | tee $O/r02n_small_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 30 -c 1 -f -o $O/r02n_step64 \
    python bench.py --workload cavity64 --graph-iters 0 --steps 40 --warmup 20 --no-cpu --no-e2e --no-extras > $O/r02n_ncu_step64.log 2>&1
