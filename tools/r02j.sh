#!/bin/bash
# round 2, GPU call J: multi-iteration kernel (parity + small boxes), generic containers on dGrid / bGrid incl. the new user
# lambda, C++ Skeleton graph on the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dense.py -x -q -m gpu > $O/r02j_pytest_dense.log 2>&1
echo "pytest rc=$?" >> $O/r02j_pytest_dense.log
tail -4 $O/r02j_pytest_dense.log
timeout 1200 python -m pytest tests/test_cpp_veneer.py -q -m gpu > $O/r02j_pytest_veneer.log 2>&1
echo "pytest rc=$?" >> $O/r02j_pytest_veneer.log
tail -6 $O/r02j_pytest_veneer.log
B="python bench.py --no-e2e --no-cpu --no-extras --steps 200 --warmup 20"
for w in cavity64 cavity96 cavity128 cavity160; do
  for v in "" "--no-persistent" "--graph-iters 50" "--graph-iters 50 --vec 2" "--vec 2" "--graph-iters 0"; do
    timeout 300 $B --workload $w $v > $O/r02j_small.json 2> $O/r02j_small.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02j_small.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"]*1000,2), "us/step", "frac", round(j["roofline"]["frac"],3), "iters/launch", j["config"]["iterations_per_launch"])
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02j_small.err").read()[-400:])
PY
  done
done 2>&1 | tee $O/r02j_small_sweep.log
timeout 600 neon_b200/cpp/bin/generic-containers --deviceIds 0 --bench 512 > $O/r02j_generic_bench.log 2>&1
timeout 600 neon_b200/cpp/bin/generic-containers --deviceIds 0 --bench 256 >> $O/r02j_generic_bench.log 2>&1
tail -4 $O/r02j_generic_bench.log
timeout 300 python bench.py --no-e2e --no-cpu --no-extras > $O/r02j_bench_default.json 2>/dev/null
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02j_bench_default.json").read().strip().splitlines()[-1]); print("default", round(j["value"]), j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["frac_of_nominal_8TBps"])
PY
