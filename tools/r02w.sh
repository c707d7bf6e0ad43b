#!/bin/bash
# round 2, GPU call W: bGrid kernel walking several blocks per CTA with the next block's info line fetched ahead
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_block.py -x -q -m gpu > $O/r02w_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02w_pytest.log
tail -3 $O/r02w_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 30 --warmup 5"
for w in sphere bcavity512; do
for v in "--rows-log2 1" "--rows-log2 2" "" "--rows-log2 8" "--rows-log2 15"; do
  $B --workload $w $v > $O/r02w_b.json 2> $O/r02w_b.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02w_b.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"],4), "ms/step", "frac", round(j["roofline"]["frac"],3))
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02w_b.err").read()[-400:])
PY
done
done 2>&1 | tee $O/r02w_block_sweep.log
