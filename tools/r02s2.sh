#!/bin/bash
# round 2, GPU call: staggered launch chain for one-wave boxes (all planes on the counters, plane z starts z * dt late)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 1000 --warmup 100 --persistent --chain-graph --chain-early 15 --graph-iters 100"
for w in cavity64 cavity48 cavity96; do
  for vec in 4 2; do
    for ns in 0 60 120 200 400 800; do
      NLBM_CHAIN_STAGGER_NS=$ns $B --workload $w --vec $vec > $O/r02s2.json 2> $O/r02s2.err
      python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02s2.json").read().strip().splitlines()[-1])
    print("$w vec $vec stagger $ns ns:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step")
except Exception as e:
    print("$w vec $vec stagger $ns FAILED", e, open("gpurun_out/r02s2.err").read()[-300:])
PY
    done
  done
done 2>&1 | tee $O/r02s2_stagger_sweep.log
