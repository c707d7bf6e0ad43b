#!/bin/bash
# round 2, final GPU call on one B200: complete -m gpu suite, smoke(), both bench arms as the driver runs them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > $O/r02final_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02final_pytest.log
tail -4 $O/r02final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02final_smoke.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02final_bench_reference.json 2> $O/r02final_bench_reference.err ) 2>&1 | grep real
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02final_bench_default.json 2> $O/r02final_bench_default.err ) 2>&1 | grep real
python - <<PY
import json
j=json.loads(open("gpurun_out/r02final_bench_default.json").read().strip().splitlines()[-1])
print("headline", round(j["value"]), j["ms_per_step"], j["roofline"]["frac"], "traffic", j["roofline"]["traffic"], "e2e", j["e2e"] and round(j["e2e"]["value"]), "arith_ref", j["arith_reference"] and round(j["arith_reference"]["value"]), "launches", j["gpu_launches"], j["clocks"])
for e in j["extra_configs"]:
    print(" ", e.get("key"), e.get("error") or (round(e["value"]), round(e["roofline"]["frac"],3), e.get("issue"), e["roofline"].get("l2_resident_copy")))
print("cpu", j["cpu_baseline"])
PY
