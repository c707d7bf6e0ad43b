#!/bin/bash
# round 2, GPU call M: launch chain (programmatic dependent launches, plane-wise waits) for small boxes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "several_iterations" > $O/r02m_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02m_pytest.log
tail -3 $O/r02m_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 400 --warmup 40"
for w in cavity32 cavity48 cavity64 cavity96 cavity128 cavity192 cavity256; do
  for v in "" "--persistent" "--persistent --chain-graph" "--persistent --vec 2" "--persistent --vec 4" "--persistent --graph-iters 50" "--persistent --chain-graph --graph-iters 50" "--persistent --kernel coop"; do
    $B --workload $w --graph-iters 10 $v > $O/r02m_small.json 2> $O/r02m_small.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02m_small.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step", "frac", round(j["roofline"]["frac"],3), "iters/launch", j["config"]["iterations_per_launch"])
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02m_small.err").read()[-400:])
PY
  done
done 2>&1 | tee $O/r02m_small_sweep.log
