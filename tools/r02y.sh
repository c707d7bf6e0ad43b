#!/bin/bash
# round 2, GPU call Y: bGrid — arithmetic neighbours for interior blocks of dense-ordered partitions (no info-line wait)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_sanitizer.py -x -q -m gpu > $O/r02y_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02y_pytest.log
tail -3 $O/r02y_pytest.log
timeout 900 python -m pytest tests/test_cpp_veneer.py tests/test_gpu_multiproc.py -x -q -m gpu -k "bgrid or generic or block" > $O/r02y_pytest_cpp.log 2>&1
echo "pytest rc=$?" >> $O/r02y_pytest_cpp.log
tail -3 $O/r02y_pytest_cpp.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-extras --steps 30 --warmup 5"
for w in sphere bcavity512; do
for v in "" "--flags-summary-first" "--opts-extra 0x10000000" "--arith reference"; do
  $B --workload $w $v > $O/r02y_b.json 2> $O/r02y_b.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02y_b.json").read().strip().splitlines()[-1])
    print("$w [$v]:", round(j["value"]), "MLUPS", round(j["ms_per_step"],4), "ms/step", "frac", round(j["roofline"]["frac"],3))
except Exception as e:
    print("$w [$v] FAILED", e, open("gpurun_out/r02y_b.err").read()[-400:])
PY
done
done 2>&1 | tee $O/r02y_block_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_block_step -s 2 -c 1 -f -o $O/r02y_block_sphere \
    python bench.py --workload sphere --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02y_ncu_block.log 2>&1
