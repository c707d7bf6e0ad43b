#!/bin/bash
# round 2, GPU call T: complete -m gpu suite (no -x), D3Q27 fp64 access-width check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > $O/r02t_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02t_pytest.log
tail -4 $O/r02t_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 30 --warmup 5"
for v in "" "--vec 2" "--vec 1 --rows-log2 3" "--vec 1 --rows-log2 2"; do
  $B --workload d3q27f64 $v > $O/r02t_q27.json 2> $O/r02t_q27.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02t_q27.json").read().strip().splitlines()[-1])
    print("d3q27f64 [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"],4), "ms/step", "frac", round(j["roofline"]["frac"],3))
except Exception as e:
    print("d3q27f64 [$v] FAILED", e, open("gpurun_out/r02t_q27.err").read()[-400:])
PY
done 2>&1 | tee $O/r02t_q27_sweep.log
