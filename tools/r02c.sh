#!/bin/bash
# round 2, GPU call C: the conversion-lean evaluation of REFERENCE arithmetic (same bits): parity, then A/B against the
# operand-for-operand transcription and FAST, ncu of the new kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dense.py tests/test_gpu_block.py -x -q -m gpu > $O/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02c_pytest.log
tail -5 $O/r02c_pytest.log
timeout 600 python -m pytest tests/test_gpu_fast_parity.py -q -m gpu > $O/r02c_fastparity.log 2>&1
tail -3 $O/r02c_fastparity.log
B="python bench.py --no-e2e --no-cpu --no-extras --steps 100 --warmup 10"
for v in "exact:--arith reference" "literal:--arith reference --opts-extra 0x40000000" "fast:" "exact_slab:--arith reference --workload slab1024" \
         "exact_256:--arith reference --workload cavity256" "exact_sphere:--arith reference --workload sphere" "fast_sphere:--workload sphere"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 $B $flags > $O/r02c_bench_$name.json 2> $O/r02c_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02c_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4), j["clocks"])
except Exception as e:
    print("$name FAILED", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_step -s 2 -c 1 -f -o $O/r02c_exact512 \
    python bench.py --arith reference --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > $O/r02c_ncu_exact512.log 2>&1
# the default bench line as the driver runs it (all extras), wall-clocked
( time timeout 900 python bench.py > $O/r02c_bench_full.json 2> $O/r02c_bench_full.err ) 2> $O/r02c_bench_full.time
tail -3 $O/r02c_bench_full.time
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02c_bench_refarm.json 2> $O/r02c_bench_refarm.err ) 2> $O/r02c_bench_refarm.time
