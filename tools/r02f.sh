#!/bin/bash
# round 2, GPU call F: the reference's own host code over the library (integration test); block-shape sweep of the step kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_reference_integration.py -q -m gpu -s > $O/r02f_integration.log 2>&1
tail -15 $O/r02f_integration.log
for n in 256 512; do
  NLBM_SHIM_ARITH=fast timeout 300 oracle/_ref/ref_lbm_b200 --device gpu --n $n --iters 110 --bench 10 --fp float 2>/dev/null | grep ref_bench | sed "s/^/shim fast $n: /"
  timeout 300 oracle/_ref/ref_lbm_b200 --device gpu --n $n --iters 110 --bench 10 --fp float 2>/dev/null | grep ref_bench | sed "s/^/shim reference-arith $n: /"
  timeout 300 oracle/_ref/ref_lbm --device gpu --n $n --iters 110 --bench 10 --fp float 2>/dev/null | grep ref_bench | sed "s/^/unmodified reference $n: /"
done > $O/r02f_shim_bench.log 2>&1
cat $O/r02f_shim_bench.log
B="python bench.py --no-e2e --no-cpu --no-extras --steps 100 --warmup 10"
for w in cavity512 slab1024 cavity256 cavity128 d3q27f64 cavity1024; do
  for r in 0 1 2 3 4; do
    timeout 300 $B --workload $w --rows-log2 $r > $O/r02f_rows_${w}_$r.json 2> $O/r02f_rows_${w}_$r.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02f_rows_${w}_$r.json").read().strip().splitlines()[-1])
    print("$w rows-log2=$r", round(j["value"]), round(j["ms_per_step"],4), round(j["roofline"]["frac"],4))
except Exception as e:
    print("$w rows-log2=$r FAILED", e)
PY
  done
done 2>&1 | tee $O/r02f_rows_sweep.log
