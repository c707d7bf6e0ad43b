#!/bin/bash
# round 2, GPU call A: parity of the new kernel variants + pipelined halo, then A/B of the step kernel at 512^3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02a_box.txt
nproc >> gpurun_out/r02a_box.txt
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_multiproc.py -x -q -m gpu > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python -m pytest tests/test_gpu_fast_parity.py -x -q -m gpu > gpurun_out/r02a_fastparity.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_fastparity.log
tail -5 gpurun_out/r02a_fastparity.log
B="python bench.py --no-e2e --no-cpu --steps 100 --warmup 10"
for v in "default:" "flagwords:--opts-extra 0x10000000" "noxfix:--opts-extra 0x20000000" "r01:--opts-extra 0x30000000" \
         "exp3:--experiment 3" "summary:--flags-summary-first" "default2:"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 $B $flags > gpurun_out/r02a_bench_$name.json 2> gpurun_out/r02a_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02a_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4), j["clocks"])
except Exception as e:
    print("$name FAILED", e)
PY
done
for w in cavity256 cavity128 cavity64 d3q27f64 slab1024; do
  timeout 300 $B --workload $w > gpurun_out/r02a_bench_$w.json 2> gpurun_out/r02a_bench_$w.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02a_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4))
except Exception as e:
    print("$w FAILED", e)
PY
done
