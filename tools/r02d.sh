#!/bin/bash
# round 2, GPU call D: exact evaluation v2 (shared-reciprocal division): self-test + parity, A/B over vector widths, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dense.py tests/test_gpu_block.py -x -q -m gpu > $O/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02d_pytest.log
tail -5 $O/r02d_pytest.log
timeout 600 python -m pytest tests/test_gpu_fast_parity.py -q -m gpu -k "19-float32-128" > $O/r02d_fastparity.log 2>&1
tail -3 $O/r02d_fastparity.log
B="python bench.py --no-e2e --no-cpu --no-extras --steps 100 --warmup 10"
for v in "exact:--arith reference" "exact_v2:--arith reference --vec 2" "exact_v1:--arith reference --vec 1" "literal:--arith reference --opts-extra 0x40000000" "fast:" \
         "exact_rows4:--arith reference --rows-log2 3" "exact_sphere:--arith reference --workload sphere" "exact_q27:--arith reference --workload d3q27f64"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 $B $flags > $O/r02d_bench_$name.json 2> $O/r02d_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02d_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4), j["clocks"])
except Exception as e:
    print("$name FAILED", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/r02d_exact512 \
    python bench.py --arith reference --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02d_ncu_exact512.log 2>&1
