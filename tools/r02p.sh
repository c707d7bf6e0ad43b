#!/bin/bash
# round 2, GPU call P: full -m gpu suite with the launch chain in place, smoke(), the default bench line, ncu of k_dense_chain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02p_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02p_pytest.log
tail -4 $O/r02p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02p_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02p_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02p_bench_default.json 2> $O/r02p_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02p_bench_reference.json 2> $O/r02p_bench_reference.err; echo "bench ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_chain -s 12 -c 1 -f -o $O/r02p_chain128 \
    python bench.py --workload cavity128 --steps 40 --warmup 20 --no-cpu --no-e2e --no-extras > $O/r02p_ncu_chain128.log 2>&1
python - <<PY
import json
j=json.loads(open("gpurun_out/r02p_bench_default.json").read().strip().splitlines()[-1])
print("headline", round(j["value"]), j["ms_per_step"], j["roofline"]["frac"], "e2e", j["e2e"] and round(j["e2e"]["value"]), "arith_ref", j["arith_reference"] and round(j["arith_reference"]["value"]))
for e in j["extra_configs"]:
    print(" ", e.get("key"), e.get("error") or (round(e["value"]), round(e["roofline"]["frac"],3), e.get("issue")))
print("cpu", j["cpu_baseline"])
PY
tail -c 700 $O/r02p_bench_reference.json
