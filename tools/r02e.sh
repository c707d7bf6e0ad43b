#!/bin/bash
# round 2, GPU call E: new FAST fp32 arithmetic (reference rounding points): long-run parity + full GPU suite; speed A/B;
# exact evaluation through the TMA-fed kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast_parity.py -q -m gpu > $O/r02e_fastparity.log 2>&1
tail -3 $O/r02e_fastparity.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_fast_parity.py > $O/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02e_pytest.log
tail -5 $O/r02e_pytest.log
B="python bench.py --no-e2e --no-cpu --no-extras --steps 100 --warmup 10"
for v in "fast:" "fast2:" "fast_sphere:--workload sphere" "fast_slab:--workload slab1024" "fast_256:--workload cavity256" "exact:--arith reference" "exact_tma:--arith reference --kernel tma" \
         "fast_tma:--kernel tma" "exact_rows4:--arith reference --rows-log2 3" "fast_rows4:--rows-log2 3" "fast_rows2:--rows-log2 2"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 $B $flags > $O/r02e_bench_$name.json 2> $O/r02e_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02e_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4), j["clocks"])
except Exception as e:
    print("$name FAILED", e)
PY
done
