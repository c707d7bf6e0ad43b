#!/bin/bash
# round 2, GPU call L: multi-iteration kernel, one block per SM without spills, own barrier
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "several_iterations" > $O/r02l_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02l_pytest.log
tail -3 $O/r02l_pytest.log
B="python bench.py --no-e2e --no-cpu --no-extras --steps 200 --warmup 20"
for w in cavity32 cavity48 cavity64 cavity80 cavity96 cavity128; do
  for v in "" "--no-persistent" "--vec 2" "--graph-iters 50"; do
    timeout 300 $B --workload $w $v > $O/r02l_small.json 2> $O/r02l_small.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02l_small.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step", "frac", round(j["roofline"]["frac"],3), "iters/launch", j["config"]["iterations_per_launch"])
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02l_small.err").read()[-400:])
PY
  done
done 2>&1 | tee $O/r02l_small_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_multi -s 2 -c 1 -f -o $O/r02l_multi64 \
    python bench.py --workload cavity64 --steps 40 --warmup 20 --no-cpu --no-e2e --no-extras > $O/r02l_ncu_multi64.log 2>&1
